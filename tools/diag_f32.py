import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle as O
from raytracercore_b200 import RAY_DT, RTC_F32, RTC_F64, Context, Scene
scene = Scene.from_file(os.path.join(ROOT, "tests", "scenes", "cornell_bounce.scene"))
scene.override(width=64, height=64, recursion=8)
ora = O.OracleScene(scene, seed=7)
rng = np.random.default_rng(1)
n = 4096
rays = np.zeros(n, RAY_DT)
rays["origin"] = rng.uniform(-1.8, 1.8, (n, 3)) * [1, 1, 0.5] + [0, 0, -1]
d = rng.normal(size=(n, 3))
rays["dir"] = d / np.linalg.norm(d, axis=1, keepdims=True)
want = ora.trace_closest(rays)
ctx = Context(0, RTC_F32)
ctx.load(scene, seed=7)
got = ctx.trace_closest(rays)
same = (got["prim"] == want["prim"]) & (want["prim"] >= 0)
rel = np.abs(got["t"] - want["t"]) / np.maximum(np.abs(want["t"]), 1e-300)
dn = np.linalg.norm(got["normal"] - want["normal"], axis=1)
for p in range(22):
    m = same & (want["prim"] == p)
    if m.any():
        print("prim %2d  n=%4d  max rel t %.3g  max dn %.3g" % (p, m.sum(), rel[m].max(), dn[m].max()))
worst = np.argsort(-np.where(same, rel, 0))[:5]
for i in worst:
    print(i, rays[i], "\n  got", got[i], "\n want", want[i])
