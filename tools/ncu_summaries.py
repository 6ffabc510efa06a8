"""Turns raw ncu output (brought back in gpurun_out/) into the small CSV summaries committed under profiles/.

  python tools/ncu_summaries.py launches gpurun_out/r1d_launches.csv profiles/r1d_launch_summary.csv
  python tools/ncu_summaries.py full gpurun_out/r1d_trace.ncu-rep profiles/r1d_trace_q8_ncu_full.csv
  python tools/ncu_summaries.py traffic soup1m/f32 gpurun_out/r2_trace.ncu-rep gpurun_out/r2_shade.ncu-rep RAYS
      RAYS = closest-hit queries of the captured pass (prof_step.py prints them); the captures hold every trace / shade launch
      of that one pass. Updates profiles/traffic.json, which bench.py reads for `roofline.traffic` and `limiters`.
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "lts__t_requests_srcunit_tex_op_read.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_config_size", "launch__grid_size", "launch__block_size",
]


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("==")) if r]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    tot = OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("rtc::", "")
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(r[iu], 1.0)
        n, t = tot.get(name, (0, 0.0))
        tot[name] = (n + 1, t + v)
    total = sum(t for _, t in tot.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list %s: per-kernel totals over all captured launches (cold-cache, serialised: compare shares)\n" % src)
        f.write("# kernel, launches, total_ms, share\n")
        for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write("%s, %d, %.3f, %.4f\n" % (k, n, t, t / total))
    print(open(dst).read())


def full(src, dst):
    out = open(src).read() if src.endswith(".csv") else subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none, %s; one column per captured launch of %s\n" % (src, data[0][ik]))
        f.write("metric, " + ", ".join("launch%d" % i for i in range(len(data))) + "\n")
        for m in FULL_METRICS:
            if m in hdr:
                i = hdr.index(m)
                f.write("%s [%s], %s\n" % (m, units[i], ", ".join(r[i] for r in data)))
    print(open(dst).read())


def _raw(src):
    # src: an .ncu-rep, or its raw page already exported on the GPU box (ncu -i x.ncu-rep --page raw --csv > x_raw.csv)
    out = open(src).read() if src.endswith(".csv") else subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]

    def col(name, scale_units=None):
        i = hdr.index(name)
        vals = []
        for r in data:
            v = float(r[i].replace(",", ""))
            if scale_units:
                v *= scale_units.get(units[i], 1.0)
            vals.append(v)
        return vals
    return col, len(data)


BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
SECS = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}


def traffic(key, trace_rep, shade_rep, rays, dst="profiles/traffic.json"):
    import json
    import os
    rays = float(rays)
    entry = {}
    col, n = _raw(trace_rep)
    t = col("gpu__time_duration.sum", SECS)
    byts = [a + b for a, b in zip(col("dram__bytes_read.sum", BYTES), col("dram__bytes_write.sum", BYTES))]
    tw = lambda name: sum(v * w for v, w in zip(col(name), t)) / sum(t)  # noqa: E731  (time-weighted mean over the launches)
    entry["trace"] = {
        "source": "%s: %d launches of one pass, %d rays" % (os.path.basename(trace_rep), n, int(rays)),
        "dram_bytes_per_ray": sum(byts) / rays, "dram_gbs": sum(byts) / sum(t) / 1e9, "launch_ms_under_ncu": [round(x * 1e3, 4) for x in t],
        "issue_active_pct": tw("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "l1_wavefront_pct": tw("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "alu_pipe_pct": tw("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "lanes_per_instruction": tw("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "l2_hit_pct": tw("lts__t_sector_hit_rate.pct"),
        "warp_instructions_per_ray": sum(col("smsp__inst_executed.sum")) / rays,
    }
    if shade_rep and shade_rep != "-":
        col, n = _raw(shade_rep)
        t = col("gpu__time_duration.sum", SECS)
        byts = [a + b for a, b in zip(col("dram__bytes_read.sum", BYTES), col("dram__bytes_write.sum", BYTES))]
        entry["shade"] = {"source": "%s: %d launches of one pass, %d shaded paths" % (os.path.basename(shade_rep), n, int(rays)),
                          "dram_bytes_per_path": sum(byts) / rays, "dram_gbs": sum(byts) / sum(t) / 1e9,
                          "launch_ms_under_ncu": [round(x * 1e3, 4) for x in t],
                          "long_scoreboard_per_issue": sum(v * w for v, w in zip(col("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"), t)) / sum(t)}
    try:
        allv = json.load(open(dst))
    except Exception:
        allv = {}
    allv[key] = entry
    json.dump(allv, open(dst, "w"), indent=1, sort_keys=True)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(*sys.argv[2:])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
