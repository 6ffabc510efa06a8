"""Parity at the sizes BASELINE.json states.

Intersection: for every BASELINE config the oracle traces a strided sample of the config's own frame (its scene at full
primitive count, its image size and recursion) and dumps every Scene.RayTrace call it makes -- camera rays and each bounce's
rays with their skip hits. The CUDA closest-hit kernel is handed exactly those (ray, skip) batches through rtc_trace_closest
and must return the oracle's primitive and inside flag (bit-exact in f64 mode; f32 mode within the documented, counted
ambiguity classes) and t / normal within 1e-5 (f64) / 1e-4 (f32) relative.

Shading: converged f32 images of the configs against committed high-spp renders of the oracle
(tests/golden/shading_*.npz, made by tests/golden/make_shading_fixtures.py): per-tile mean radiance within a Monte Carlo
confidence bound and whole-image RMSE <= 1 % of the mean.
"""
import os
import time

import numpy as np
import pytest

import oracle as O
from conftest import SCENES
from parity import check_hits
from raytracercore_b200 import RTC_F32, RTC_F64, Context, Scene

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# BASELINE.json configs: scene at full size, image size and recursion as stated; `grid` = pixels sampled per axis
CONFIGS = {
    "c1_bounce": dict(file="cornell_bounce.scene", width=512, height=512, recursion=8, grid=128),
    "c2_die": dict(file="die.scene", width=1920, height=1080, recursion=3, grid=128),
    "c3_soup1m": dict(synth="soup", n=1_000_000, sseed=0xC3, jitter=0.01, width=2048, height=2048, recursion=4, grid=128),
    "c4_spheres100k": dict(synth="spheres", n=100_000, sseed=0xC4, jitter=0.0, width=1920, height=1080, recursion=8, grid=128),
    "c5_soup10m": dict(synth="soup", n=10_000_000, sseed=0xC5, jitter=0.004, width=3840, height=2160, recursion=4, grid=96),
}
# f32-mode caps per bounce class = 10 x the counts measured on the B200 (fractions of the batch; parity.py logs the counts)
CAPS = {
    # 3562 of 40 829 bounce rays: the reference's re-hit of the sphere the ray leaves, 1e-13..1e-9 along the ray (the class the
    # shading test below documents); not 10 x but 1.4 x the measured count -- the class is a property of the scene
    "c4_spheres100k": {"bounce": dict(max_unresolvable_frac=0.12)},
}


def make_scene(cfg):
    if "synth" in cfg:
        sc = Scene.synthetic(cfg["synth"], cfg["n"], cfg["sseed"], cfg["jitter"])
    else:
        sc = Scene.from_file(os.path.join(SCENES, cfg["file"]))
    sc.override(width=cfg["width"], height=cfg["height"], recursion=cfg["recursion"])
    return sc


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_intersection_parity_on_dumped_path_batches(name):
    cfg = CONFIGS[name]
    sc = make_scene(cfg)
    ora = O.OracleScene(sc, seed=31)
    g = cfg["grid"]
    xs = (np.arange(g) * cfg["width"]) // g + (cfg["width"] // g) // 2
    ys = (np.arange(g) * cfg["height"]) // g + (cfg["height"] // g) // 2
    xy = np.stack(np.meshgrid(xs, ys), -1).reshape(-1, 2).astype(np.int32)
    rays, skip, want, bounce = ora.dump_path_rays(xy, np.full(len(xy), 5, np.uint32))
    n_b = np.bincount(bounce, minlength=cfg["recursion"] + 1)
    assert n_b[0] == len(xy) and len(rays) > len(xy) and n_b[1:].sum() > 0  # camera rays and bounces present
    assert (skip["prim"][bounce == 0] == -1).all() and (skip["prim"][bounce > 0] >= 0).all()
    hit_frac = (want["prim"] >= 0).mean()
    assert hit_frac > 0.2, hit_frac
    modes = [(RTC_F64, 1e-5, True), (RTC_F32, 1e-4, False)]
    arrays = sc.arrays()
    for prec, tol, exact in modes:
        ctx = Context(0, prec)
        ctx.upload_scene(sc)
        ctx.upload_bvh(*sc.bvh())
        got = ctx.trace_closest(rays, skip)
        ctx.close()
        # tiny spheres (r ~ 0.01): the normal is (P - C) / r, so f32 coordinates limit it to ~1e-4 / r relative
        ntol = 2e-3 if (not exact and cfg.get("synth") == "spheres") else None
        for cls, sel in (("camera", bounce == 0), ("bounce", bounce > 0)):
            cp = CAPS.get(name, {}).get(cls, {})
            check_hits(got[sel], want[sel], tol, exact, origins=rays["origin"][sel], dirs=rays["dir"][sel], normal_tol=ntol,
                       label="%s/%s/%s" % (name, "f64" if exact else "f32", cls), skip=skip[sel], arrays=arrays, **cp)
    ora.close()


# ---- shading parity against committed high-spp oracle renders --------------------------------------------
FIXTURES = {
    # fixture: (scene config, samples per pixel of the GPU's config-sized pass, of its converged pass)
    "c1": (dict(file="cornell_bounce.scene"), 64, 4096),
    "c2": (dict(file="die.scene"), 64, 4096),
    "c3": (dict(synth="soup", n=1_000_000, sseed=0xC3, jitter=0.01), 16, 2048),
    "c4": (dict(synth="spheres", n=100_000, sseed=0xC4, jitter=0.0), 16, 2048),
    "c4f": (dict(synth="spheres", n=100_000, sseed=0xC4, jitter=0.0), 16, 2048),
}
# (fixture, arithmetic mode). The production f32 mode is held to the reference's image on C1, C2 and C3. On C4 (10^5 spheres of
# radius 0.004..0.012 seen from 3.5 units away) the reference's own image is decided by f64 rounding noise: its sphere hit
# points lie 1e-13..1e-6 off the surface, and Util.NearEnough = 1e-24 accepts the re-hit of the same sphere that follows
# 1e-11 further on for 8.5 % of all bounce rays -- those paths stay trapped inside the sphere until the recursion limit and
# return its (mostly zero) emission. RTC_F64 restates that arithmetic and reproduces the image path by path ("c4", f64);
# no f32 arithmetic can (the deciding distances are below its resolution), so the f32 mode is checked against the oracle
# run with the f32 mode's documented self-hit rule ("c4f", oracle/rtc_oracle.h: orc_set_selfhit_mode) -- everything but
# that one deviation. Measured on the B200: f32 mean radiance = 1.18 x the reference's on this scene.
SHADING_CASES = [("c1", RTC_F32), ("c2", RTC_F32), ("c3", RTC_F32), ("c4f", RTC_F32), ("c4", RTC_F64)]


def tiles(a, t):
    h, w = a.shape[:2]
    return a.reshape((h // t, t, w // t, t) + a.shape[2:]).sum(axis=(1, 3))


LUM = np.array([0.299, 0.587, 0.114])


@pytest.mark.parametrize("name,prec", SHADING_CASES)
def test_shading_parity_against_high_spp_oracle_render(name, prec):
    path = os.path.join(GOLDEN, "shading_%s.npz" % name)
    fx = np.load(path)
    base, spp_cfg, spp_conv = FIXTURES[name]
    W, H, T = int(fx["width"]), int(fx["height"]), int(fx["tile"])
    cfg = dict(base, width=W, height=H, recursion=int(fx["recursion"]))
    assert int(fx["spp"]) >= 16 * spp_cfg  # the oracle render has >= 16 x the samples of the config-sized GPU pass
    sc = make_scene(cfg)
    ctx = Context(0, prec)
    o_spp = int(fx["spp"])
    # f32: streams independent of the oracle's (seed in the fixture). f64 is the replay mode: it is given the oracle's own
    # seed, so its converged pass below re-draws the very samples of the fixture (path for path) and the
    # comparison carries no Monte Carlo noise of its own; its config-sized chunks use sample indices beyond the fixture's.
    replay = prec == RTC_F64
    ctx.load(sc, seed=int(fx["seed"]) if replay else 977)
    first = o_spp if replay else 0
    if replay:
        spp_conv = o_spp
    # K independent chunks of the config-sized pass give the pass's own per-tile variance
    K = 8
    chunk = []
    for k in range(K):
        ctx.clear_accum()
        ctx.render(first + k * spp_cfg, spp_cfg)
        rgb, s, m = ctx.read_accum()
        assert np.all(s + m == spp_cfg)
        chunk.append((rgb, s.astype(np.int64), m.astype(np.int64)))
    ctx.clear_accum()
    ctx.render(0 if replay else K * spp_cfg, spp_conv)
    c_rgb, c_s, c_m = ctx.read_accum()
    assert np.all(c_s.astype(np.int64) + c_m == spp_conv)
    ctx.close()

    o_rgb, o_s, o_m = fx["rgb"], fx["samples"], fx["misses"]
    # (1) silhouettes: camera-ray misses depend on geometry and pixel jitter only -- per-tile miss fraction within binomial noise
    g_miss = tiles(c_m.astype(np.int64), T) / float(T * T * spp_conv)
    o_miss = o_m / float(T * T * o_spp)
    var = np.maximum(o_miss * (1 - o_miss), 1e-4) * (1.0 / (T * T * spp_conv) + 1.0 / (T * T * o_spp))
    zz = np.abs(g_miss - o_miss) / np.sqrt(var)
    if replay:  # the same pixel jitter as the oracle: the miss counts are the oracle's, tile for tile
        assert np.array_equal(tiles(c_m.astype(np.int64), T), o_m), np.abs(tiles(c_m.astype(np.int64), T) - o_m).max()
    assert np.mean(zz <= 4.5) >= 0.995 and abs(g_miss.mean() - o_miss.mean()) <= 5e-4, (zz.max(), g_miss.mean(), o_miss.mean())

    # radiance per pixel sample = colour sum / (samples + misses) (a miss contributes nothing), luminance, 32x32-pixel tiles
    def tile_lum(rgb_t, spp):
        return (rgb_t @ LUM) / float(TT * TT * spp)

    TT = 32
    f = TT // T
    agg = lambda a: tiles(a, f)
    ref = tile_lum(agg(o_rgb), o_spp)
    # (2) the config-sized pass: per-tile mean within a Monte Carlo confidence bound; the variance of a tile mean comes from
    # the K independent chunks (sample variance of the chunk means / 1) plus the oracle render's own (its two halves)
    cm = np.stack([tile_lum(tiles(c[0], TT), spp_cfg) for c in chunk])  # [K, th, tw]
    var_g = cm.var(axis=0, ddof=1)
    half0 = tile_lum(agg(fx["rgb_half0"]), o_spp / 2.0)
    half1 = tile_lum(agg(o_rgb - fx["rgb_half0"]), o_spp / 2.0)
    var_o = ((half0 - half1) ** 2) / 4.0
    floor = (1e-3 * np.maximum(ref, ref.mean() * 1e-2)) ** 2
    z = np.abs(cm - ref[None]) / np.sqrt(var_g[None] + var_o[None] + floor[None])
    assert np.mean(z <= 4.0) >= 0.99, (name, float(np.mean(z <= 4.0)), float(z.max()))
    # and the mean of the K chunks (K x the samples) must tighten accordingly: no bias hiding under the single-pass noise
    zk = np.abs(cm.mean(axis=0) - ref) / np.sqrt(var_g / K + var_o + floor)
    assert np.mean(zk <= 4.0) >= 0.98, (name, float(np.mean(zk <= 4.0)), float(zk.max()))
    # (3) converged image: whole-image RMSE of the tile means within 1 % of the mean radiance (north_star), no noise allowance
    conv = tile_lum(tiles(c_rgb, TT), spp_conv)
    rmse = float(np.sqrt(np.mean((conv - ref) ** 2)))
    print("shading[%s]: rmse/mean = %.4f, mean %.5f vs %.5f" % (name, rmse / ref.mean(), conv.mean(), ref.mean()))
    assert rmse <= (0.002 if replay else 0.01) * ref.mean(), (name, rmse, ref.mean())
    assert abs(conv.mean() - ref.mean()) <= 0.004 * ref.mean()
    # per channel as well (tints): image means within 0.5 %
    for ch in range(3):
        a = tiles(c_rgb, TT)[..., ch].sum() / spp_conv
        b = agg(o_rgb)[..., ch].sum() / o_spp
        assert abs(a - b) <= 0.005 * max(b, 1e-9), (name, ch, a, b)
