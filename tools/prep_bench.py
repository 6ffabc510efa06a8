"""Scene.Prepare timing: host path (binned SAH + flatten on the host threads) against rtc_prepare_device, per workload.
python tools/prep_bench.py [n_prims] [repeats] [--no-host]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raytracercore_b200 import RTC_BUILDER_PLOC, RTC_BUILDER_SAH, RTC_F32, Context, Scene  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sc = Scene.synthetic("soup", n, 0xC3, 0.01)
if "--no-host" not in sys.argv:
    t = time.time()
    host = Context(0, RTC_F32)
    host.upload_scene(sc)
    t1 = time.time()
    host.build_bvh()
    host.sync()
    print("host: upload_scene %.3f s, build + flatten %.3f s" % (t1 - t, time.time() - t1), flush=True)
    host.close()
dev = Context(0, RTC_F32)
for r in range(reps):
    t = time.time()
    dev.upload_scene(sc)
    t1 = time.time()
    st = dev.prepare_device(RTC_BUILDER_PLOC if "--ploc" in sys.argv else RTC_BUILDER_SAH)
    dev.sync()
    print("device run %d: upload_scene %.1f ms, prepare %.1f ms (boxes %.1f, build %.1f / %d levels, flatten %.1f); %d wide nodes, depth %d"
          % (r, (t1 - t) * 1e3, (time.time() - t1) * 1e3, st.boxes_ms, st.build_ms, st.build_levels, st.flatten_ms, st.n_wide_nodes, st.wide_depth), flush=True)
