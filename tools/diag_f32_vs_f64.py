"""Diagnostic: where do f32-mode paths leave the f64 (oracle-exact) paths? Same seed in both modes, per-path radiance compared,
first diverging bounce of a sample of differing paths classified from the debug traces."""
import sys, os, collections
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytracercore_b200 import Scene, Context, RTC_F32, RTC_F64
from raytracercore_b200 import _native as N

name = sys.argv[1] if len(sys.argv) > 1 else "spheres"
if name == "spheres":
    sc = Scene.synthetic("spheres", 100_000, 0xC4, 0.0); sc.override(width=256, height=128, recursion=8)
elif name == "soup":
    sc = Scene.synthetic("soup", 1_000_000, 0xC3, 0.01); sc.override(width=256, height=256, recursion=4)
else:
    sc = Scene.from_file(os.path.join(ROOT, "tests", "scenes", name)); sc.override(width=256, height=256, recursion=8)
c32, c64 = Context(0, RTC_F32), Context(0, RTC_F64)
c32.load(sc, seed=5); c64.load(sc, seed=5)
W, H = sc.width, sc.height
tot32 = np.zeros(3); tot64 = np.zeros(3); ndiff = 0; npaths = 0
cls = collections.Counter()
examples = []
for s in range(16):
    a, b = c32.render_samples(s), c64.render_samples(s)
    ma, mb = np.all(a == -1, axis=2), np.all(b == -1, axis=2)
    a0, b0 = np.where(ma[..., None], 0, a), np.where(mb[..., None], 0, b)
    tot32 += a0.sum(axis=(0, 1)); tot64 += b0.sum(axis=(0, 1))
    d = ~np.isclose(a0, b0, rtol=1e-3, atol=1e-4).all(axis=2)
    ndiff += int(d.sum()); npaths += W * H
    ys, xs = np.nonzero(d)
    for k in range(min(len(ys), 40)):
        x, y = int(xs[k]), int(ys[k])
        ta, tb = c32.debug_trace(x, y, s), c64.debug_trace(x, y, s)
        j = 0
        while j < min(len(ta), len(tb)) and ta[j].type == tb[j].type and ta[j].hit.prim == tb[j].hit.prim and ta[j].hit.inside == tb[j].hit.inside:
            j += 1
        if j < min(len(ta), len(tb)):
            key = ("b%d" % j, "type %s vs %s" % (N.BOUNCE_TYPES[ta[j].type], N.BOUNCE_TYPES[tb[j].type]),
                   "prim %s" % ("same" if ta[j].hit.prim == tb[j].hit.prim else "diff"),
                   "inside %s" % ("same" if ta[j].hit.inside == tb[j].hit.inside else "diff"),
                   "prevtype %s" % (N.BOUNCE_TYPES[tb[j - 1].type] if j else "-"))
            if len(examples) < 12:
                examples.append((x, y, s, j, [(N.BOUNCE_TYPES[r.type], r.hit.prim, r.hit.inside, r.hit.t) for r in ta],
                                 [(N.BOUNCE_TYPES[r.type], r.hit.prim, r.hit.inside, r.hit.t) for r in tb]))
        else:
            key = ("len %d vs %d" % (len(ta), len(tb)),)
        cls[key] += 1
print("paths", npaths, "differ", ndiff, "frac %.4f" % (ndiff / npaths))
print("mean radiance f32", tot32 / npaths, "f64", tot64 / npaths, "ratio", tot32 / np.maximum(tot64, 1e-30))
for k, v in cls.most_common(25):
    print(v, k)
for e in examples:
    print(e)
