#!/usr/bin/env python
"""bench.py — headline benchmark of the render hot path (BASELINE.json): Mrays/s (closest-hit queries per second)
and spp·Mpix/s on a named scene, 1..8 B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload soup1m|soup10m|spheres100k|die|bounce]
  python bench.py --impl reference ...     # the reference's CPU path (oracle port) on the host cores

A step is one progressive pass of the wavefront path tracer: --spp samples for every pixel of the image on every
rank (weak scaling: rank r renders the sample range [r*spp, (r+1)*spp) of the frame), followed for N > 1 by the single
per-frame NCCL reduce of the accumulation buffer to rank 0. `value` times the passes with everything resident in
HBM; `e2e` re-does the same through the C ABI with host buffers: scene + BVH hand-over from host memory, render,
read-back of the SampleSet planes.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# spp = samples per pixel per step and GPU, chosen so that one step is one wavefront of up to 32 Mi paths (the pool's cap):
# ~65 ms per step on the 1 M-triangle scene, the order of the reference's 100 ms status / bitmap cadence (FullRaytracer.cs:33,369).
WORKLOADS = {
    # name: (description, builder kwargs)
    "soup1m": dict(desc="synthetic 1M-triangle random soup, 2048x2048, recursion 4 (BASELINE C3)", synth="soup", n=1_000_000,
                   seed=0xC3, jitter=0.01, width=2048, height=2048, recursion=4, spp=8),
    "soup10m": dict(desc="synthetic 10M-triangle random soup, 3840x2160, recursion 4 (BASELINE C5)", synth="soup", n=10_000_000,
                    seed=0xC5, jitter=0.004, width=3840, height=2160, recursion=4, spp=4),
    "spheres100k": dict(desc="synthetic 100k spheres mirror/glass/diffuse, 1920x1080, recursion 8 (BASELINE C4)", synth="spheres",
                        n=100_000, seed=0xC4, jitter=0.0, width=1920, height=1080, recursion=8, spp=16),
    "die": dict(desc="die scene 1920x1080, recursion 3, DOF (BASELINE C2)", file="die.scene", width=1920, height=1080, recursion=3, spp=16),
    "bounce": dict(desc="Cornell 'bounce' scene 512x512, recursion 8 (BASELINE C1)", file="cornell_bounce.scene", width=512, height=512,
                   recursion=8, spp=128),
}


def make_scene(wl, small=False):
    from raytracercore_b200 import Scene
    w = WORKLOADS[wl]
    if "synth" in w:
        sc = Scene.synthetic(w["synth"], w["n"], w["seed"], w["jitter"])
    else:
        sc = Scene.from_file(os.path.join(ROOT, "tests", "scenes", w["file"]))
    sc.override(width=w["width"], height=w["height"], recursion=w["recursion"])
    return sc


def algorithmic_bytes_per_ray(n_prims):
    """ray in 32 + hit out 32 + one root-to-leaf path of 64-B nodes + one 48-B primitive."""
    return 112 + 64 * max(1, math.ceil(math.log2(max(2, n_prims))))


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.15 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return None
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for k, nm in enumerate(names):
                if r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons), "samples": len(rows)}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (the oracle port: collect-all-leaves BVH walk, f64) on the host
    cores, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as O
    wl = WORKLOADS[args.workload]
    sc = make_scene(args.workload)
    t0 = time.time()
    ora = O.OracleScene(sc, seed=1)
    threads = os.cpu_count() or 1
    W, H = wl["width"], wl["height"]
    # bounded sample: a centred square crop, 1 spp per step, sized from a probe so one step is a few seconds
    probe = 32
    x0, y0 = (W - probe) // 2, (H - probe) // 2
    t = time.time()
    _, _, _, rays = ora.render(0, 1, rect=(x0, y0, x0 + probe, y0 + probe), threads=threads)
    dt = max(time.time() - t, 1e-4)
    rate = rays / dt
    side = int(min(min(W, H), max(probe, math.sqrt(max(1.0, args.ref_seconds * rate / max(rays / (probe * probe), 1e-9))))))
    side = max(probe, (side // 8) * 8)
    x0, y0 = (W - side) // 2, (H - side) // 2
    rect = (x0, y0, x0 + side, y0 + side)
    for i in range(args.warmup):
        ora.render(i, 1, rect=rect, threads=threads)
    tot_rays = 0
    t = time.time()
    for i in range(args.steps):
        _, _, _, r = ora.render(args.warmup + i, 1, rect=rect, threads=threads)
        tot_rays += r
    dt = time.time() - t
    mrays = tot_rays / dt / 1e6
    paths = side * side * args.steps
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": mrays, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "description": wl["desc"], "spp_mpix_per_s": paths / dt / 1e6},
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": threads, "kind": "port",
                         "sample": "%dx%d centre crop of the %dx%d frame, 1 spp per step, %d steps" % (side, side, W, H, args.steps)},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "setup_s": time.time() - t0,
    }
    print(json.dumps(line))
    return 0


def cpu_baseline(args, sc, seconds):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as O
    wl = WORKLOADS[args.workload]
    ora = O.OracleScene(sc, seed=1)
    threads = os.cpu_count() or 1
    W, H = wl["width"], wl["height"]
    probe = 32
    x0, y0 = (W - probe) // 2, (H - probe) // 2
    t = time.time()
    _, _, _, rays = ora.render(0, 1, rect=(x0, y0, x0 + probe, y0 + probe), threads=threads)
    dt = max(time.time() - t, 1e-4)
    per_px = dt / (probe * probe)
    side = int(min(min(W, H), max(probe, math.sqrt(seconds / per_px))))
    side = max(probe, (side // 8) * 8)
    x0, y0 = (W - side) // 2, (H - side) // 2
    t = time.time()
    _, _, _, rays = ora.render(1, 1, rect=(x0, y0, x0 + side, y0 + side), threads=threads)
    dt = time.time() - t
    ora.close()
    return {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
            "sample": "%dx%d centre crop of the %dx%d frame, 1 spp (%.1f s of CPU work, C++ f64 restatement of the reference walk)" % (side, side, W, H, dt),
            "spp_mpix_per_s": side * side / dt / 1e6}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="soup1m", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="samples per pixel per step (default: per workload)")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-seconds", type=float, default=6.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--max-paths", type=int, default=0)
    ap.add_argument("--no-other-scenes", action="store_true", help="skip the short per-scene runs of the other BASELINE configs")
    ap.add_argument("--with-soup10m", action="store_true", help="include the 10 M-triangle scene in the per-scene runs (its BVH build takes ~10 s)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    from raytracercore_b200 import RTC_F32, RTC_F64, RTC_OPT_KERNEL_TIMING, RTC_OPT_MAX_PATHS, Context
    from raytracercore_b200 import _native as N

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: librtcore_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = WORKLOADS[args.workload]
    spp = args.spp or wl["spp"]
    W, H = wl["width"], wl["height"]

    t_setup = time.time()
    sc = make_scene(args.workload)
    nodes, n_nodes, root = sc.bvh()  # Scene.Prepare: host BVH build (cached between renders, like the reference)
    t_bvh = time.time() - t_setup
    prec = RTC_F64 if args.precision == "f64" else RTC_F32
    ctx = Context(local, prec)
    stream = torch.cuda.Stream()  # a real (non-default) stream: handle 0 would mean "the context's own stream"
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    if args.max_paths:
        ctx.set_option(RTC_OPT_MAX_PATHS, args.max_paths)
    ctx.load(sc, seed=1)
    uid = None
    if world > 1:
        obj = [Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        uid = obj[0]
        ctx.comm_init(world, rank, uid)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def frame(step):
        # one frame: every rank adds its sample range, then the single per-frame collective
        if world > 1:
            ctx.clear_accum()
        ctx.render((step * world + rank) * spp, spp)
        if world > 1:
            ctx.reduce_accum(0)

    # ---------------- device-resident timing ----------------
    for i in range(args.warmup):
        frame(i)
    barrier()
    ctx.reset_stats()
    ctx.set_option(RTC_OPT_KERNEL_TIMING, 1)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record(stream)
    for i in range(args.steps):
        frame(args.warmup + i)
    e1.record(stream)
    barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    st = ctx.stats()
    ctx.set_option(RTC_OPT_KERNEL_TIMING, 0)
    tms = torch.tensor([ms, float(st.rays), float(st.paths)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = tms.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tms.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_max, rays_all, paths_all = float(mx[0]), float(sm[1]), float(sm[2])
    else:
        ms_max, rays_all, paths_all = ms, float(st.rays), float(st.paths)
    value = rays_all / (ms_max * 1e-3) / 1e6

    # ---------------- end-to-end through the C ABI with host buffers ----------------
    # One step = what FullRaytracer.Start does per frame with host-resident inputs: the host's prepared scene image
    # (Scene.Prepare output: device-layout primitives + BVH in pinned host memory) is copied to the device, camera and
    # parameters are handed over as host structs, the pass is rendered, the per-frame collective runs (N > 1) and the
    # SampleSet planes are read back into pinned host memory.
    e2e = None
    if not args.no_e2e:
        baked = ctx.bake()
        h2d = baked.nbytes + 8 * 40  # + rtc_camera / rtc_params structs
        d2h = W * H * (24 + 4 + 4)
        par = sc.params(1)
        cam = sc.camera()
        e2e_steps = max(2, min(args.steps, 4))
        pin_rgb = torch.empty((H, W, 3), dtype=torch.float64, pin_memory=True)
        pin_s = torch.empty((H, W), dtype=torch.int32, pin_memory=True)
        pin_m = torch.empty((H, W), dtype=torch.int32, pin_memory=True)
        out = (pin_rgb.data_ptr(), pin_s.data_ptr(), pin_m.data_ptr())

        def e2e_step(step):
            ctx.upload_baked(baked)
            ctx.set_params(par)
            ctx.set_camera(cam)
            ctx.clear_accum()
            if world == 1:
                # one call: render + read-back, each band's rows streaming out while the next band renders
                ctx.render_read(step * spp, spp, out)
                return
            ctx.render((step * world + rank) * spp, spp)
            ctx.reduce_accum(0)
            if rank == 0:
                ctx.read_accum(out)
            else:
                ctx.sync()

        e2e_step(0)
        barrier()
        ctx.reset_stats()
        tt = time.time()
        for i in range(e2e_steps):
            e2e_step(1 + i)
        barrier()
        dt = time.time() - tt
        st2 = ctx.stats()
        v = torch.tensor([dt, float(st2.rays)], dtype=torch.float64, device="cuda")
        if world > 1:
            mx = v.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm2 = v.clone()
            dist.all_reduce(sm2, op=dist.ReduceOp.SUM)
            dt, rays2 = float(mx[0]), float(sm2[1])
        else:
            rays2 = float(st2.rays)
        e2e = {"value": rays2 / dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": dt / e2e_steps * 1e3,
               "what": "per step: H2D of the baked scene image from pinned host memory (shading half behind the first trace "
                       "launch), camera+params structs, render, reduce (N>1), D2H of the SampleSet planes to pinned host "
                       "memory (N=1: rtc_render_read, band read-back overlapped with the next band); wall clock, max over ranks"}

    if rank == 0:
        n_prims = sc.n_prims
        bpr = algorithmic_bytes_per_ray(n_prims)
        trace_ms = st.ms[N.RTC_K_TRACE]
        trace_launches = st.launches[N.RTC_K_TRACE]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = (float(st.rays) * bpr / 1e9) / (trace_ms * 1e-3) if trace_ms > 0 else None
        # dram__bytes_read.sum + dram__bytes_write.sum of k_trace_q8 from the `ncu --set full` capture in
        # profiles/r1d_trace_q8_ncu_full.csv (soup1m, f32): 364.2 / 495.7 / 553.2 MB for launches of 4 194 304 / 3 933 350 /
        # 2 529 048 rays = 132.6 B per ray, scaled to this run's rays per launch; null for workloads without a capture
        traffic = 132.6 * float(st.rays) / max(1, trace_launches) if (args.workload == "soup1m" and args.precision == "f32") else None
        roofline = {"bound": "hbm", "kernel": "k_trace_q8" if args.precision == "f32" else "k_trace<double>", "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                    "traffic_source": "profiles/r1d_trace_q8_ncu_full.csv: 132.6 B of DRAM traffic per ray x rays_per_launch (L2-resident tree: traffic << algorithmic bytes)" if traffic else None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                    "algorithmic_bytes_per_ray": bpr, "rays_per_launch": float(st.rays) / max(1, trace_launches),
                    "avg_launch_ms": trace_ms / max(1, trace_launches),
                    "kernel_ms": {N.KERNEL_NAMES[k]: st.ms[k] for k in range(N.RTC_K_COUNT)},
                    "kernel_share_of_step": {N.KERNEL_NAMES[k]: st.ms[k] / ms for k in range(N.RTC_K_COUNT)}}
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": args.workload, "description": wl["desc"], "width": W, "height": H, "spp_per_step_per_gpu": spp,
                       "recursion": wl["recursion"], "n_prims": n_prims, "partition": "sample ranges per rank, scene replicated, 1 NCCL reduce per frame" if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (scene %.0f MB + path pool; no flush)" % ((n_prims * (64 + 48 + 64)) / 1e6),
                       "spp_mpix_per_s": paths_all / (ms_max * 1e-3) / 1e6, "rays_per_path": rays_all / max(1.0, paths_all),
                       "bvh_build_s": t_bvh},
            "roofline": roofline,
            "gpu_launches": int(sum(st.launches[k] for k in range(N.RTC_K_COUNT))),
            "clocks": clocks,
        }
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, sc, args.cpu_seconds)
        if world == 1 and not args.no_other_scenes:
            # BASELINE.json quotes the metric "per scene": short device-resident runs of the other configs (3 warm-up + 4
            # timed steps each, CUDA events on the launching stream), reported beside the headline workload
            others = [w for w in ("bounce", "die", "spheres100k") + (("soup10m",) if args.with_soup10m else ()) if w != args.workload]
            per_scene = {}
            for w in others:
                wl2 = WORKLOADS[w]
                sc2 = make_scene(w)
                c2 = Context(local, prec)
                c2.set_stream(stream.cuda_stream)
                c2.load(sc2, seed=1)
                for i in range(3):
                    c2.render(i * wl2["spp"], wl2["spp"])
                c2.sync()
                c2.reset_stats()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record(stream)
                for i in range(4):
                    c2.render((3 + i) * wl2["spp"], wl2["spp"])
                a1.record(stream)
                c2.sync()
                t2 = a0.elapsed_time(a1) * 1e-3
                st3 = c2.stats()
                per_scene[w] = {"value": st3.rays / t2 / 1e6, "unit": "Mrays/s", "spp_mpix_per_s": st3.paths / t2 / 1e6,
                                "ms_per_step": t2 / 4 * 1e3, "description": wl2["desc"], "spp_per_step": wl2["spp"],
                                "n_prims": sc2.n_prims}
                c2.close()
            line["per_scene"] = per_scene
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
