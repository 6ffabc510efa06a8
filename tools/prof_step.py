"""Profiling helper: one workload, a few passes, per-kernel event times and traversal counters.
Used under ncu on the GPU box (see profiles/README.md)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from raytracercore_b200 import RTC_F32, RTC_F64, RTC_OPT_COUNTERS, RTC_OPT_KERNEL_TIMING, RTC_OPT_MAX_PATHS, RTC_OPT_WAVES, Context  # noqa: E402
from raytracercore_b200 import _native as N  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="soup1m")
ap.add_argument("--spp", type=int, default=1)
ap.add_argument("--passes", type=int, default=2)
ap.add_argument("--size", type=int, default=0, help="override width=height")
ap.add_argument("--precision", default="f32")
ap.add_argument("--counters", action="store_true")
ap.add_argument("--max-paths", type=int, default=0)
ap.add_argument("--waves", type=int, default=0, help="wavefronts in flight (RTC_OPT_WAVES); implies no per-kernel timing")
ap.add_argument("--reorder", type=int, default=0, help="RTC_OPT_REORDER mode")
ap.add_argument("--device-bvh", type=int, default=-1, help="build the tree on the GPU with this search radius (0 = default)")
a = ap.parse_args()
sc = bench.make_scene(a.workload)
if a.size:
    sc.override(width=a.size, height=a.size)
ctx = Context(0, RTC_F64 if a.precision == "f64" else RTC_F32)
if a.max_paths:
    ctx.set_option(RTC_OPT_MAX_PATHS, a.max_paths)
t_build = time.time()
if a.device_bvh >= 0:
    ctx.upload_scene(sc)
    t_build = time.time()
    rounds = ctx.build_bvh(device=True, radius=a.device_bvh)
    print("device BVH build + flatten + upload: %.3f s (%d rounds, radius %d)" % (time.time() - t_build, rounds, a.device_bvh))
    ctx.set_params(sc.params(1))
    ctx.set_camera(sc.camera(None))
else:
    nodes, n, root = sc.bvh()
    print("host BVH build: %.3f s" % (time.time() - t_build))
    t_build = time.time()
    ctx.upload_scene(sc)
    ctx.upload_bvh(nodes, n, root)
    print("flatten + upload: %.3f s" % (time.time() - t_build))
    ctx.set_params(sc.params(1))
    ctx.set_camera(sc.camera(None))
if a.reorder:
    from raytracercore_b200 import RTC_OPT_REORDER
    ctx.set_option(RTC_OPT_REORDER, a.reorder)
ctx.render(0, a.spp)
ctx.sync()
ctx.reset_stats()
if a.waves:
    ctx.set_option(RTC_OPT_WAVES, a.waves)
else:
    ctx.set_option(RTC_OPT_KERNEL_TIMING, 1)
if a.counters:
    ctx.set_option(RTC_OPT_COUNTERS, 1)
t = time.time()
for i in range(a.passes):
    ctx.render((1 + i) * a.spp, a.spp)
ctx.sync()
dt = time.time() - t
st = ctx.stats()
print("wall %.3f s  paths %d rays %d  -> %.1f Mrays/s (wall)" % (dt, st.paths, st.rays, st.rays / dt / 1e6))
for k in range(N.RTC_K_COUNT):
    print("  %-10s launches %4d  ms %9.3f" % (N.KERNEL_NAMES[k], st.launches[k], st.ms[k]))
if st.ms[N.RTC_K_TRACE] > 0:
    print("trace-only Mrays/s: %.1f" % (st.rays / st.ms[N.RTC_K_TRACE] / 1e3))
if a.counters:
    print("nodes/ray %.1f prims/ray %.2f" % (st.nodes_visited / st.rays, st.prims_tested / st.rays))
    if st.node_steps:
        print("lanes per node step %.1f, per leaf step %.1f; warp node steps %d leaf steps %d" % (
            st.nodes_visited / st.node_steps, st.prims_tested / max(1, st.leaf_steps), st.node_steps, st.leaf_steps))
