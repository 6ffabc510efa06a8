"""Turns raw ncu output (brought back in gpurun_out/) into the small CSV summaries committed under profiles/.

  python tools/ncu_summaries.py launches gpurun_out/r1d_launches.csv profiles/r1d_launch_summary.csv
  python tools/ncu_summaries.py full gpurun_out/r1d_trace.ncu-rep profiles/r1d_trace_q8_ncu_full.csv
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
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
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


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
