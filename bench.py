#!/usr/bin/env python
"""bench.py — headline benchmark of the render hot path (BASELINE.json): Mrays/s (closest-hit queries per second)
and spp·Mpix/s on a named scene, 1..8 B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload soup1m|soup10m|spheres100k|die|bounce]
  python bench.py --impl reference ...     # the reference's CPU path (oracle port) on the host cores

A step is one progressive pass of the wavefront path tracer: --spp samples for every pixel of the image on every
rank (weak scaling: rank r renders the sample range [r*spp, (r+1)*spp) of the frame), followed for N > 1 by the single
per-frame NCCL reduce of the accumulation buffer to rank 0. `value` times the passes with everything resident in
HBM; `e2e` re-does the same through the C ABI with host buffers: scene + BVH hand-over from host memory (N > 1: one
PCIe upload on rank 0, then rtc_bcast_scene over NVLink), render, reduce, read-back of the SampleSet planes.

Besides the headline workload (BASELINE C3, the configuration the north-star ratio is quoted on) the one JSON line carries
`c5`: the 10 M-triangle scene at 3840x2160 (BASELINE C5, the configuration BASELINE quotes at 1/2/4/8 GPUs) in a weak- and a
strong-scaling form, and `per_scene`: short runs of the other BASELINE configs at the same N.
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
C5_STRONG_TOTAL_SPP = 8  # strong-scaling form of C5: this many samples per pixel per frame, split over the ranks


def make_scene(wl, small=False):
    from raytracercore_b200 import Scene
    w = WORKLOADS[wl]
    if "synth" in w:
        sc = Scene.synthetic(w["synth"], w["n"], w["seed"], w["jitter"])
    else:
        sc = Scene.from_file(os.path.join(ROOT, "tests", "scenes", w["file"]))
    sc.override(width=w["width"], height=w["height"], recursion=w["recursion"])
    return sc


def n_prims_of(wl):
    """Primitive count of a workload without building it (the file scenes are small: build those)."""
    w = WORKLOADS[wl]
    return w["n"] if "synth" in w else make_scene(wl).n_prims


def static_config(wl, spp, world, n_prims):
    """The workload's configuration: the same dictionary on the GPU arm and on the reference arm."""
    w = WORKLOADS[wl]
    scene_mb = n_prims * (96 / 4.4 + 48 + 32 + 16 + 12) / 1e6  # f32 image: ~1 node per 4.4 primitives, records, materials, ids
    pool_mb = w["width"] * w["height"] * spp * 104 / 1e6       # path state written and re-read every bounce
    l2 = ("inputs larger than L2: scene image %.0f MB + %.0f MB of path state per step (no flush)" % (scene_mb, pool_mb)) if scene_mb > 126 else (
        "%.0f MB of path state per step streams through the 126 MB L2 (the %.1f MB scene image stays resident; no flush)" % (pool_mb, scene_mb))
    return {"workload": wl, "description": w["desc"], "width": w["width"], "height": w["height"], "spp_per_step_per_gpu": spp,
            "recursion": w["recursion"], "n_prims": n_prims,
            "partition": "sample ranges per rank, scene on rank 0 + NVLink broadcast, 1 NCCL reduce per frame" if world > 1 else "single GPU",
            "l2": l2}


def algorithmic_bytes_per_ray(n_prims):
    """ray in 32 + hit out 32 + one root-to-leaf path of 64-B nodes + one 48-B primitive."""
    return 112 + 64 * max(1, math.ceil(math.log2(max(2, n_prims))))


def load_traffic():
    """profiles/traffic.json (written by tools/ncu_summaries.py traffic from `ncu --set full` captures): measured DRAM bytes per
    ray / per shaded path and the limiter percentages, per workload and kernel."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return {}


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


# ---------------------------------------------------------------------------------------------------------------
# CPU legs: the oracle port of the reference walk on the host cores, on a lattice of the whole frame
# ---------------------------------------------------------------------------------------------------------------
def lattice_for(ora, W, H, threads, seconds):
    """Stride of a pixel lattice over the whole frame whose 1-spp render takes about `seconds` on this host: probed on a
    coarse lattice first. The lattice keeps the frame's mix of rays (silhouettes, background, all depths)."""
    probe = max(1, int(math.sqrt(W * H / 2048.0)))
    t = time.time()
    _, _, _, rays = ora.render_lattice((probe, probe), (probe // 2, probe // 2), 0, 1, threads=threads)
    dt = max(time.time() - t, 1e-4)
    n_probe = len(range(probe // 2, W, probe)) * len(range(probe // 2, H, probe))
    per_px = dt / max(1, n_probe)
    want = max(n_probe, min(W * H, seconds / per_px))
    stride = max(1, int(math.sqrt(W * H / want)))
    return stride


def cpu_sample(ora, W, H, threads, stride, first, n):
    t = time.time()
    _, _, _, rays = ora.render_lattice((stride, stride), (stride // 2, stride // 2), first, n, threads=threads)
    dt = time.time() - t
    px = len(range(stride // 2, W, stride)) * len(range(stride // 2, H, stride))
    return rays, px * n, dt


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (the oracle port: collect-all-leaves BVH walk, f64) on the host
    cores with all their threads, on a bounded lattice sample of the same frame. The process loads the scene half of the
    host mirror (librtcore_host.so) and the oracle only: no CUDA library is mapped."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    os.environ["RTC_B200_HOST_ONLY"] = "1"
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as O
    wl = WORKLOADS[args.workload]
    spp = args.spp or wl["spp"]
    t0 = time.time()
    sc = make_scene(args.workload)
    ora = O.OracleScene(sc, seed=1)
    threads = os.cpu_count() or 1
    W, H = wl["width"], wl["height"]
    stride = lattice_for(ora, W, H, threads, args.ref_seconds)
    for i in range(args.warmup):
        cpu_sample(ora, W, H, threads, stride, i, 1)
    tot_rays = tot_paths = 0
    t = time.time()
    for i in range(args.steps):
        r, p, _ = cpu_sample(ora, W, H, threads, stride, args.warmup + i, 1)
        tot_rays += r
        tot_paths += p
    dt = time.time() - t
    mrays = tot_rays / dt / 1e6
    sample = "lattice of every %d-th pixel in x and y over the whole %dx%d frame (%d pixels), 1 spp per step, %d steps" % (
        stride, W, H, tot_paths // max(1, args.steps), args.steps)
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": mrays, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": static_config(args.workload, spp, max(1, args.gpus), sc.n_prims),
        "measured": {"spp_mpix_per_s": tot_paths / dt / 1e6, "rays_per_path": tot_rays / max(1, tot_paths)},
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": threads, "per_core": mrays / threads, "kind": "port", "sample": sample},
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
    stride = lattice_for(ora, W, H, threads, seconds)
    rays, paths, dt = cpu_sample(ora, W, H, threads, stride, 1, 1)
    ora.close()
    v = rays / dt / 1e6
    return {"value": v, "unit": "Mrays/s", "cores": threads, "per_core": v / threads, "kind": "port",
            "sample": "lattice of every %d-th pixel in x and y over the whole %dx%d frame (%d pixels), 1 spp (%.1f s of CPU work, C++ f64 "
                      "restatement of the reference walk)" % (stride, W, H, paths, dt),
            "spp_mpix_per_s": paths / dt / 1e6}


# ---------------------------------------------------------------------------------------------------------------
# GPU legs
# ---------------------------------------------------------------------------------------------------------------
class Job:
    """One workload on this rank's GPU: scene prepared on rank 0 (Scene.Prepare: host BVH build + bake), handed to the
    other ranks over NVLink, one context per rank."""

    def __init__(self, env, workload, prec, max_paths=0):
        import torch.distributed as dist
        from raytracercore_b200 import RTC_OPT_MAX_PATHS, Context
        self.env, self.workload = env, workload
        self.wl = WORKLOADS[workload]
        self.W, self.H = self.wl["width"], self.wl["height"]
        rank, world = env["rank"], env["world"]
        self.ctx = Context(env["local"], prec)
        self.ctx.set_stream(env["stream"].cuda_stream)
        if max_paths:
            self.ctx.set_option(RTC_OPT_MAX_PATHS, max_paths)
        self.sc = None
        t = time.time()
        meta = [None]
        self.prepare = None
        if rank == 0:
            from raytracercore_b200 import RTC_BUILDER_SAH
            self.sc = make_scene(workload)
            self.sc.desc()  # (the flattened description, made once by the host scene: not part of Scene.Prepare)
            if os.environ.get("RTC_BENCH_HOST_PREPARE") == "1":  # the round-1 path: host binned SAH, host flatten, H2D of the image
                t = time.time()
                self.sc.bvh()
                self.t_bvh = time.time() - t
                t = time.time()
                self.ctx.load(self.sc, seed=1)
                self.ctx.sync()
                self.t_flatten = time.time() - t
                self.prepare = {"where": "host", "total_s": self.t_bvh + self.t_flatten}
            else:
                # Scene.Prepare on the device (rtc_prepare_device): the host SAH tree, node for node, and the device layout, byte for
                # byte, without the tree or the image ever crossing PCIe (tests/test_gpu_prepare.py)
                t = time.time()
                self.ctx.load(self.sc, seed=1, device_prepare=RTC_BUILDER_SAH)
                self.ctx.sync()
                total = time.time() - t
                st = self.ctx.prepare_stats
                self.t_bvh = (st.boxes_ms + st.build_ms) * 1e-3
                self.t_flatten = total - self.t_bvh  # rtc_upload_scene's copy of the description + collapse / quantisation / records
                self.prepare = {"where": "device", "total_s": total, "boxes_ms": st.boxes_ms, "build_ms": st.build_ms, "flatten_ms": st.flatten_ms,
                                "build_levels": st.build_levels, "wide_nodes": st.n_wide_nodes, "wide_depth": st.wide_depth}
            self.par, self.cam = self.sc.params(1), self.sc.camera()
            meta = [dict(par=bytes(self.par), cam=bytes(self.cam), n_prims=self.sc.n_prims, t_bvh=self.t_bvh, t_flatten=self.t_flatten)]
        self.t_bcast = 0.0
        if world > 1:
            from raytracercore_b200 import _native as N
            dist.broadcast_object_list(meta, src=0)
            m = meta[0]
            uid = [Context.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            self.ctx.comm_init(world, rank, uid[0])
            env["barrier"]()
            t = time.time()
            self.ctx.bcast_scene(0)
            self.ctx.sync()
            env["barrier"]()
            self.t_bcast = time.time() - t
            if rank != 0:
                self.par, self.cam = N.Params.from_buffer_copy(m["par"]), N.Camera.from_buffer_copy(m["cam"])
                self.ctx.set_params(self.par)
                self.ctx.set_camera(self.cam)
                self.t_bvh, self.t_flatten = m["t_bvh"], m["t_flatten"]
            self.n_prims = m["n_prims"]
        else:
            self.n_prims = self.sc.n_prims

    def frame(self, step, spp):
        """One frame: every rank adds its sample range, then the single per-frame collective."""
        rank, world = self.env["rank"], self.env["world"]
        self.ctx.render((step * world + rank) * spp, spp)
        if world > 1:
            self.ctx.reduce_accum(0)

    def timed(self, spp, steps, warmup, kernel_timing=False, sample_clocks=False):
        """`steps` frames after `warmup`, device-resident, CUDA events on the launching stream, max over ranks."""
        import torch
        import torch.distributed as dist
        from raytracercore_b200 import RTC_OPT_KERNEL_TIMING
        env = self.env
        rank, world, stream = env["rank"], env["world"], env["stream"]
        self.ctx.clear_accum()
        for i in range(warmup):
            self.frame(i, spp)
        env["barrier"]()
        self.ctx.reset_stats()
        if kernel_timing:
            self.ctx.set_option(RTC_OPT_KERNEL_TIMING, 1)
        sampler = ClockSampler(env["local"]) if (sample_clocks and rank == 0) else None
        if sampler:
            sampler.start()
        env["barrier"]()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(stream)
        for i in range(steps):
            self.frame(warmup + i, spp)
        e1.record(stream)
        env["barrier"]()
        t1 = time.time()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop(t0, t1) if sampler else None
        st = self.ctx.stats()
        if kernel_timing:
            self.ctx.set_option(RTC_OPT_KERNEL_TIMING, 0)
        tms = torch.tensor([ms, float(st.rays), float(st.paths)], dtype=torch.float64, device="cuda")
        if world > 1:
            mx = tms.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = tms.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            ms_max, rays_all, paths_all = float(mx[0]), float(sm[1]), float(sm[2])
        else:
            ms_max, rays_all, paths_all = ms, float(st.rays), float(st.paths)
        return dict(value=rays_all / (ms_max * 1e-3) / 1e6, ms=ms_max, ms_rank=ms, rays=rays_all, paths=paths_all, stats=st, clocks=clocks,
                    ms_per_step=ms_max / steps)

    def e2e(self, spp, steps, upload_every_step=True):
        """The same frames through the C ABI with host buffers. One step = what FullRaytracer.Start does per frame with
        host-resident inputs: the prepared scene image (Scene.Prepare's output: device-layout primitives + tree in pinned host
        memory) goes to the device -- N > 1: over PCIe to rank 0 only, then to the other ranks over NVLink (rtc_bcast_scene) --,
        camera and parameters are handed over as host structs, the pass is rendered, the per-frame collective runs (N > 1) and
        the SampleSet planes are read back into pinned host memory. upload_every_step=False times one FullRaytracer.Start()
        instead: the scene goes up once (the reference prepares and caches it once per Start, Scene.cs:39-49), then `steps`
        progressive passes follow, each with its parameter hand-over, collective and read-back; the upload stays inside the
        timed region and its bytes are spread over the steps in h2d_bytes_per_step."""
        import torch
        import torch.distributed as dist
        env = self.env
        rank, world = env["rank"], env["world"]
        ctx, W, H = self.ctx, self.W, self.H
        baked = ctx.bake() if rank == 0 else None
        h2d = (baked.nbytes if baked else 0) + 8 * 40  # + rtc_camera / rtc_params structs
        d2h = W * H * (24 + 4 + 4) if rank == 0 else 0
        pin_rgb = torch.empty((H, W, 3), dtype=torch.float64, pin_memory=True) if rank == 0 else None
        pin_s = torch.empty((H, W), dtype=torch.int32, pin_memory=True) if rank == 0 else None
        pin_m = torch.empty((H, W), dtype=torch.int32, pin_memory=True) if rank == 0 else None
        out = (pin_rgb.data_ptr(), pin_s.data_ptr(), pin_m.data_ptr()) if rank == 0 else None

        def step(i, upload=True, clear=True):
            if upload:
                if rank == 0:
                    ctx.upload_baked(baked)
                if world > 1:
                    ctx.bcast_scene(0)
            ctx.set_params(self.par)
            ctx.set_camera(self.cam)
            if clear:
                ctx.clear_accum()
            if world == 1:
                # one call: render + read-back, each band's rows streaming out while the next band renders
                ctx.render_read(i * spp, spp, out)
                return
            ctx.render((i * world + rank) * spp, spp)
            ctx.reduce_accum(0)
            if rank == 0:
                ctx.read_accum(out)
            else:
                ctx.sync()

        step(0)
        env["barrier"]()
        ctx.reset_stats()
        tt = time.time()
        for i in range(steps):
            first = upload_every_step or i == 0
            step(1 + i, upload=first, clear=first)
        env["barrier"]()
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
        if baked:
            baked.close()
        if not upload_every_step:
            return {"value": rays2 / dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int((h2d - 320) / steps + 320), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": dt / steps * 1e3, "steps": steps,
                    "what": "one Start() of %d progressive passes: H2D of the baked scene image once (N>1: on rank 0, then rtc_bcast_scene over "
                            "NVLink), then per pass camera+params structs, render, reduce (N>1), D2H of the SampleSet planes to pinned host "
                            "memory; wall clock over the whole Start(), max over ranks" % steps}
        return {"value": rays2 / dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": dt / steps * 1e3,
                "what": "per step: H2D of the baked scene image from pinned host memory (shading half behind the first trace launch; "
                        "N>1: on rank 0 only, then rtc_bcast_scene = ncclBroadcast over NVLink to the other ranks), camera+params structs, "
                        "render, reduce (N>1), D2H of the SampleSet planes to pinned host memory (N=1: rtc_render_read, band read-back "
                        "overlapped with the next band); wall clock, max over ranks"}

    def close(self):
        self.ctx.close()
        self.sc = None


def roofline_blocks(job, res, precision, step_ms):
    """The contract's roofline object for the dominant kernel (algorithmic bytes / CUDA-event time of its launches against
    the measured HBM copy bandwidth) and, beside it, `limiters`: what the ncu captures name as the units that bound the
    kernels (issue slots and the L1 wavefront pipe for the traversal, DRAM bytes for shading)."""
    from raytracercore_b200 import _native as N
    st = res["stats"]
    bpr = algorithmic_bytes_per_ray(job.n_prims)
    trace_ms, shade_ms = st.ms[N.RTC_K_TRACE], st.ms[N.RTC_K_SHADE]
    trace_launches, shade_launches = st.launches[N.RTC_K_TRACE], st.launches[N.RTC_K_SHADE]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    rays = float(st.rays)
    achieved = (rays * bpr / 1e9) / (trace_ms * 1e-3) if trace_ms > 0 else None
    tr = load_traffic().get("%s/%s" % (job.workload, precision), {})
    tq = tr.get("trace")
    traffic = tq["dram_bytes_per_ray"] * rays / max(1, trace_launches) if tq else None
    roofline = {"bound": "hbm", "kernel": "k_trace_q8" if precision == "f32" else "k_trace<double>", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "traffic_source": ("profiles/traffic.json (%s): %.1f B of DRAM traffic per ray x rays_per_launch (L2-resident tree: traffic << "
                                   "algorithmic bytes)" % (tq["source"], tq["dram_bytes_per_ray"])) if tq else None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_ray": bpr, "rays_per_launch": rays / max(1, trace_launches),
                "avg_launch_ms": trace_ms / max(1, trace_launches),
                "kernel_ms": {N.KERNEL_NAMES[k]: st.ms[k] for k in range(N.RTC_K_COUNT)},
                "kernel_share_of_step": {N.KERNEL_NAMES[k]: st.ms[k] / step_ms for k in range(N.RTC_K_COUNT)}}
    limiters = None
    if tq:
        limiters = {"source": "ncu --set full captures summarised in profiles/traffic.json (not live: a number taken under a profiler is no bench value)",
                    "k_trace": {"bound": "issue slots + L1 wavefront pipe", "issue_active_pct": tq.get("issue_active_pct"),
                                "l1_wavefront_pct": tq.get("l1_wavefront_pct"), "alu_pipe_pct": tq.get("alu_pipe_pct"),
                                "lanes_per_instruction": tq.get("lanes_per_instruction"), "l2_hit_pct": tq.get("l2_hit_pct"),
                                "dram_gbs": tq.get("dram_gbs"), "dram_frac_of_peak": (tq["dram_gbs"] / peak) if tq.get("dram_gbs") else None}}
        ts = tr.get("shade")
        if ts and shade_ms > 0:
            paths_shaded = rays  # one shade per traced ray
            live = ts["dram_bytes_per_path"] * paths_shaded / 1e9 / (shade_ms * 1e-3)
            limiters["k_shade"] = {"bound": "hbm (gather latency)", "dram_bytes_per_path": ts["dram_bytes_per_path"], "achieved_gbs": live,
                                   "frac_of_peak": live / peak, "avg_launch_ms": shade_ms / max(1, shade_launches),
                                   "what": "measured DRAM bytes per shaded path (ncu) x paths shaded in the timed region / CUDA-event time of the shade launches"}
    return roofline, limiters


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
    ap.add_argument("--no-c5", action="store_true", help="skip the 10 M-triangle scene (its host BVH build takes ~10 s)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from raytracercore_b200 import RTC_F32, RTC_F64
    from raytracercore_b200 import _native as N
    from raytracercore_b200.partition import check_reduced_frame

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: librtcore_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()  # a real (non-default) stream: handle 0 would mean "the context's own stream"
    torch.cuda.set_stream(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    env = dict(rank=rank, world=world, local=local, stream=stream, barrier=barrier)
    prec = RTC_F64 if args.precision == "f64" else RTC_F32
    wl = WORKLOADS[args.workload]
    spp = args.spp or wl["spp"]

    # ---------------- the headline workload ----------------
    job = Job(env, args.workload, prec, args.max_paths)
    res = job.timed(spp, args.steps, args.warmup, sample_clocks=True)
    # per-kernel CUDA-event times come from a second, shorter run: with event timing on, the library keeps one wavefront in
    # flight (the launches of a step back to back on one stream), so that a launch's duration is its own; the timed region
    # above runs two wavefronts on two streams, whose launches overlap
    kt = job.timed(spp, max(2, min(4, args.steps)), 3, kernel_timing=True)
    parity = None
    if world > 1:
        # N-rank frame == the single-GPU frame of the same sample ranges (counters exact, sums to rounding): checked on rank 0
        parity = check_reduced_frame(job, spp)
    e2e = job.e2e(spp, max(2, min(args.steps, 4))) if not args.no_e2e else None
    line = None
    if rank == 0:
        roofline, limiters = roofline_blocks(job, kt, args.precision, kt["ms_rank"])
        roofline["timing_run"] = {"steps": kt["stats"].launches[N.RTC_K_ACCUMULATE] and max(2, min(4, args.steps)), "ms_per_step": kt["ms_per_step"],
                                  "value": kt["value"], "what": "one wavefront in flight (RTC_OPT_KERNEL_TIMING): launches back to back on one stream"}
        st = res["stats"]
        line = {
            "metric": "Mrays/s", "value": res["value"], "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": static_config(args.workload, spp, world, job.n_prims),
            "measured": {"spp_mpix_per_s": res["paths"] / (res["ms"] * 1e-3) / 1e6, "rays_per_path": res["rays"] / max(1.0, res["paths"]),
                         "bvh_build_s": job.t_bvh, "flatten_upload_s": job.t_flatten, "scene_bcast_s": job.t_bcast,
                         "prepare": job.prepare},
            "roofline": roofline,
            "gpu_launches": int(sum(st.launches[k] for k in range(N.RTC_K_COUNT))),
            "clocks": res["clocks"],
        }
        if limiters:
            line["limiters"] = limiters
        if parity is not None:
            line["multi_gpu_parity"] = parity
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, job.sc, args.cpu_seconds)
    job.close()
    del job

    # ---------------- BASELINE C5 at N ranks: weak and strong scaling ----------------
    if not args.no_c5 and args.workload != "soup10m":
        w5 = WORKLOADS["soup10m"]
        j5 = Job(env, "soup10m", prec)
        weak = j5.timed(w5["spp"], 4, 3)
        weak_e2e = j5.e2e(w5["spp"], 4, upload_every_step=False) if not args.no_e2e else None
        strong = strong_e2e = None
        if C5_STRONG_TOTAL_SPP % world == 0:
            s_spp = C5_STRONG_TOTAL_SPP // world
            strong = j5.timed(s_spp, 4, 3)
            strong_e2e = j5.e2e(s_spp, 4, upload_every_step=False) if not args.no_e2e else None
        if rank == 0:
            def pack(r, e, spp_rank, scaling):
                d = {"value": r["value"], "unit": "Mrays/s", "ms_per_step": r["ms_per_step"], "spp_per_step_per_gpu": spp_rank,
                     "spp_mpix_per_s": r["paths"] / (r["ms"] * 1e-3) / 1e6, "scaling": scaling, "steps": 4, "warmup": 3}
                if e:
                    d["e2e"] = e
                return d
            c5 = {"config": static_config("soup10m", w5["spp"], world, j5.n_prims), "n_gpus": world,
                  "bvh_build_s": j5.t_bvh, "flatten_upload_s": j5.t_flatten, "scene_bcast_s": j5.t_bcast, "prepare": j5.prepare,
                  "weak": pack(weak, weak_e2e, w5["spp"], "weak")}
            if strong:
                c5["strong"] = pack(strong, strong_e2e, C5_STRONG_TOTAL_SPP // world, "strong")
                c5["strong"]["total_spp_per_step"] = C5_STRONG_TOTAL_SPP
            line["c5"] = c5
        j5.close()
        del j5

    # ---------------- the other BASELINE configs, at N ranks like everything else ----------------
    if not args.no_other_scenes:
        # BASELINE.json quotes the metric "per scene at 1/2/4/8": short device-resident runs of the other configs (3 warm-up + 4
        # timed steps each, CUDA events on the launching stream, max over ranks; N > 1: weak scaling -- every rank adds its own
        # sample range of every frame, one reduce per frame), reported beside the headline workload
        per_scene = {}
        for w in ("bounce", "die", "spheres100k", "soup1m"):
            if w == args.workload:
                continue
            wl2 = WORKLOADS[w]
            j2 = Job(env, w, prec)
            r2 = j2.timed(wl2["spp"], 4, 3)
            per_scene[w] = {"value": r2["value"], "unit": "Mrays/s", "spp_mpix_per_s": r2["paths"] / (r2["ms"] * 1e-3) / 1e6,
                            "ms_per_step": r2["ms_per_step"], "description": wl2["desc"], "spp_per_step_per_gpu": wl2["spp"], "n_prims": j2.n_prims,
                            "n_gpus": world, "scaling": "weak"}
            j2.close()
            del j2
        if rank == 0:
            line["per_scene"] = per_scene
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
