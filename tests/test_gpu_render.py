"""Render-loop parity (north_star test 2 and the pieces around it): ray generation, per-path replay in f64 mode,
Monte-Carlo agreement of converged images in f32 mode, accumulation / tonemap / resume, debug traces, the
FullRaytracer mirror, and size-independent properties at BASELINE sizes."""
import os
import threading
import time

import numpy as np
import pytest

import oracle as O
from conftest import SCENES
from raytracercore_b200 import RAY_DT, RTC_OPT_COUNTERS, RTC_F32, RTC_F64, Context, FullRaytracer, Scene
from raytracercore_b200 import _native as N

pytestmark = pytest.mark.gpu


def cornell(w=96, h=96, rec=8):
    sc = Scene.from_file(os.path.join(SCENES, "cornell_bounce.scene"))
    sc.override(width=w, height=h, recursion=rec)
    return sc


def die(w=96, h=64, rec=3):
    sc = Scene.from_file(os.path.join(SCENES, "die.scene"))
    sc.override(width=w, height=h, recursion=rec)
    return sc


@pytest.mark.parametrize("make,cam", [(cornell, 0), (cornell, 4), (die, 0)])
def test_camera_rays_match_the_oracle(make, cam):
    sc = make()
    sc.override(camera=cam)
    ora = O.OracleScene(sc, seed=11)
    rng = np.random.default_rng(0)
    n = 20000
    xy = np.stack([rng.integers(0, sc.width, n), rng.integers(0, sc.height, n)], 1).astype(np.int32)
    smp = rng.integers(0, 1 << 20, n).astype(np.uint32)
    want = ora.camera_rays(xy, smp)
    for prec, tol in ((RTC_F64, 1e-12), (RTC_F32, 2e-4)):  # f32 draws the top 24 bits of the same uniforms; DOF amplifies them
        ctx = Context(0, prec)
        ctx.set_params(sc.params(11))
        ctx.set_camera(sc.camera())
        got = ctx.camera_rays(xy, smp)
        assert np.allclose(got["origin"], want["origin"], rtol=0, atol=tol * 10), prec
        assert np.allclose(got["dir"], want["dir"], rtol=0, atol=tol), prec
        ctx.close()


def test_orthographic_camera_rays():
    sc = Scene.from_string("size 40 30\northographic 1 2 -5  1 2 0  0 1 0  3\nsphere 0 0 0 1\n")
    ora = O.OracleScene(sc, seed=2)
    ys, xs = np.mgrid[0:30, 0:40]
    xy = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.int32)
    smp = np.full(len(xy), 3, np.uint32)
    want = ora.camera_rays(xy, smp)
    ctx = Context(0, RTC_F64)
    ctx.set_params(sc.params(2))
    ctx.set_camera(sc.camera())
    got = ctx.camera_rays(xy, smp)
    assert np.allclose(got["origin"], want["origin"], atol=1e-12) and np.allclose(got["dir"], want["dir"], atol=1e-15)
    ctx.close()


@pytest.mark.parametrize("make", [cornell, die])
def test_f64_path_replay_equals_the_oracle(make):
    """Same Philox streams, same arithmetic: per-path radiance must be the oracle's, except where last-bit libm
    differences (pow/acos/sin/cos) flip a lobe choice."""
    sc = make()
    ora = O.OracleScene(sc, seed=4)
    ctx = Context(0, RTC_F64)
    ctx.load(sc, seed=4)
    for s in (0, 7):
        want = ora.render_samples(s)
        got = ctx.render_samples(s)
        same = np.isclose(got, want, rtol=1e-9, atol=1e-12).all(axis=2)
        assert same.mean() >= 0.998, same.mean()
        assert np.array_equal(np.all(got == -1, axis=2), np.all(want == -1, axis=2))  # misses are decided by geometry alone
    ctx.close()


def test_debug_trace_matches_the_oracle():
    sc = cornell(64, 64, 8)
    ora = O.OracleScene(sc, seed=6)
    ctx = Context(0, RTC_F64)
    ctx.load(sc, seed=6)
    agree = 0
    pix = [(x, y) for x in range(4, 64, 12) for y in range(4, 64, 12)]
    for x, y in pix:
        a, b = ctx.debug_trace(x, y, 1), ora.debug_trace(x, y, 1)
        if [r.type for r in a] == [r.type for r in b]:
            agree += 1
            for ra, rb in zip(a, b):
                assert ra.hit.prim == rb.hit.prim and ra.hit.inside == rb.hit.inside
                if ra.hit.prim >= 0:
                    assert ra.hit.t == pytest.approx(rb.hit.t, rel=1e-9)
                if not np.isnan(rb.fresnel_ratio):
                    assert ra.fresnel_ratio == pytest.approx(rb.fresnel_ratio, rel=1e-9)
    assert agree >= len(pix) - 1
    ctx.close()


@pytest.mark.parametrize("make,spp", [(cornell, 192), (die, 256)])
def test_f32_images_agree_within_monte_carlo_bounds(make, spp):
    """Shading parity: per-tile mean radiance within 4 sigma of the oracle's (variance estimated from both sides'
    independent sample sets) and whole-image RMSE of the means within 1% of a higher-spp oracle render's mean."""
    sc = make(64, 48)
    ora = O.OracleScene(sc, seed=21)
    ctx = Context(0, RTC_F32)
    ctx.load(sc, seed=1234)  # different seed: statistically independent of the oracle's samples
    ctx.render(0, spp)
    g_rgb, g_s, g_m = ctx.read_accum()
    o_rgb, o_s, o_m, _ = ora.render(0, spp)
    assert np.all(g_s + g_m == spp)
    # primary misses depend on geometry + jitter only: miss fractions must agree within binomial noise
    pm_g, pm_o = g_m.sum() / (spp * g_m.size), o_m.sum() / (spp * o_m.size)
    assert abs(pm_g - pm_o) < 4 * np.sqrt(max(pm_o, 1e-3) / (spp * g_m.size)) + 1e-4
    # two independent oracle halves give the per-tile noise level
    a_rgb, a_s, _, _ = ora.render(1000, spp)
    lum = lambda rgb, s: (rgb @ [0.299, 0.587, 0.114]) / np.maximum(s, 1)
    Lg, Lo, La = lum(g_rgb, g_s), lum(o_rgb, o_s), lum(a_rgb, a_s)
    th, tw = 8, 8
    tile = lambda L: L.reshape(L.shape[0] // th, th, L.shape[1] // tw, tw).mean(axis=(1, 3))
    Tg, To, Ta = tile(Lg), tile(Lo), tile(La)
    sigma = np.abs(To - Ta) / np.sqrt(2) + 1e-3 * np.maximum(To, 1e-3)  # one-sample estimate, floored
    noise = np.sqrt(np.mean((To - Ta) ** 2) / 2)
    z = np.abs(Tg - 0.5 * (To + Ta))
    assert np.mean(z <= 4 * noise + 4 * sigma) >= 0.98
    ref = 0.5 * (To + Ta)
    rmse = np.sqrt(np.mean((Tg - ref) ** 2))
    assert rmse <= max(0.01 * ref.mean() + 3 * noise, 1e-6), (rmse, ref.mean(), noise)
    # and the image means agree to well under a percent plus noise
    assert abs(Lg.mean() - 0.5 * (Lo.mean() + La.mean())) <= 0.01 * Lo.mean() + 4 * abs(Lo.mean() - La.mean()) + 1e-6
    ctx.close()


def test_accumulation_is_deterministic_additive_and_band_invariant():
    sc = cornell(80, 56, 6)
    ctx = Context(0, RTC_F32)
    ctx.load(sc, seed=3)
    ctx.render(0, 6)
    a = ctx.read_accum()
    ctx.clear_accum()
    ctx.render(0, 2)
    ctx.render(2, 4)  # progressive passes add up exactly (sample order per pixel is fixed)
    b = ctx.read_accum()
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    ctx.clear_accum()
    ctx.set_option(N.RTC_OPT_MAX_PATHS, 1024)  # forces many small bands
    ctx.render(0, 6)
    c = ctx.read_accum()
    assert all(np.array_equal(x, y) for x, y in zip(a, c))
    # rectangles: two halves == whole
    ctx.clear_accum()
    ctx.render(0, 6, rect=(0, 0, 37, 56))
    ctx.render(0, 6, rect=(37, 0, 80, 56))
    d = ctx.read_accum()
    assert all(np.array_equal(x, y) for x, y in zip(a, d))
    # resume: write the planes into a fresh context and continue
    ctx2 = Context(0, RTC_F32)
    ctx2.load(sc, seed=3)
    ctx2.write_accum(*a)
    ctx2.render(6, 2)
    ctx.clear_accum()
    ctx.set_option(N.RTC_OPT_MAX_PATHS, 1 << 20)
    ctx.render(0, 8)
    assert all(np.array_equal(x, y) for x, y in zip(ctx.read_accum(), ctx2.read_accum()))
    ctx.close()
    ctx2.close()


def test_f64_accumulation_equals_the_oracle_where_paths_agree():
    sc = die(48, 32)
    ora = O.OracleScene(sc, seed=8)
    ctx = Context(0, RTC_F64)
    ctx.load(sc, seed=8)
    ctx.render(0, 4)
    g_rgb, g_s, g_m = ctx.read_accum()
    o_rgb, o_s, o_m, _ = ora.render(0, 4)
    assert np.array_equal(g_s, o_s) and np.array_equal(g_m, o_m)
    assert np.isclose(g_rgb, o_rgb, rtol=1e-9, atol=1e-12).all(axis=2).mean() > 0.99


def test_tonemap_is_bit_exact():
    sc = cornell(64, 64, 5)
    ctx = Context(0, RTC_F32)
    ctx.load(sc, seed=5)
    ctx.render(0, 3)
    rgb, s, m = ctx.read_accum()
    for exposure, back, ba in ((1.0, (0, 0, 0), 0.0), (2.5, (.2, .4, .9), 1.0), (0.3, (1, 1, 1), 0.5)):
        got = ctx.tonemap(exposure, back, ba)
        want = O.tonemap(rgb, s, m, exposure, back, ba)
        diff = np.abs((got.view(np.uint8).astype(int) - want.view(np.uint8).astype(int)))
        assert diff.max() <= 1 and (diff > 0).mean() < 1e-3  # pow() last-bit differences may move a channel by one code
    # the golden tonemap vectors through the device
    import json
    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))
    p = sc.params(1)
    p.width, p.height = 1, 1
    ctx.set_params(p)
    for v in kat["tonemap"]:
        ctx.write_accum(np.array([[v["rgb"]]], float), np.array([[v["samples"]]], np.uint32), np.array([[v["misses"]]], np.uint32))
        assert int(ctx.tonemap(v["exposure"], v["back"], v["back_a"])[0, 0]) == v["argb"]
    ctx.close()


def test_ambient_miss_and_debug_geom():
    sc = cornell(48, 48, 4)
    sc.set_ambient((-1, -1, -1))  # `ambient miss`: bounced misses count as Misses (SceneLoader.cs:183-188)
    ora = O.OracleScene(sc, seed=2)
    ctx = Context(0, RTC_F64)
    ctx.load(sc, seed=2)
    ctx.render(0, 4)
    _, s, m = ctx.read_accum()
    _, os_, om, _ = ora.render(0, 4)
    assert np.array_equal(s, os_) and np.array_equal(m, om)
    sc.set_ambient((0, 0, 0))
    sc.set_debug_geom(True)
    ora = O.OracleScene(sc, seed=2)
    ctx.load(sc, seed=2)
    assert np.array_equal(ctx.render_samples(0), ora.render_samples(0))  # no transcendental on this path: bit-exact
    ctx.close()


def test_full_raytracer_mirror():
    sc = cornell(64, 48, 5)
    log = []
    rt = FullRaytracer(sc, device=0, precision=RTC_F32, seed=1, update_status=lambda r, text, progress: log.append((text, progress)))
    assert rt.GetBitmap() is None and rt.GetSampleSet(3, 3) == ((0.0, 0.0, 0.0), 0, 0) and not rt.IsRunning
    rt.Exposure = 1.5
    rt.Start(samples_per_pass=2, max_samples=6)  # blocking, like FullRaytracer.Start
    assert not rt.IsRunning and log[0][0] == "Preparing scene..." and log[1][0] == "Beginning render..."
    assert len(log) == 5 and log[-1][0].startswith("Tiles: 3 Elapsed: ") and " 6.00/px " in log[-1][0] and log[-1][0].endswith("/px/sec")
    assert log[-1][1] == pytest.approx(6 / 1006)
    rgb, s, m = rt.GetSampleSet(32, 24)
    assert s + m == 6
    assert rt.GetSampleSet(10 ** 6, -5)[1:] == rt.GetSampleSet(63, 0)[1:]  # clamped like FullRaytracer.cs:137-138
    bmp = rt.GetBitmap()
    assert bmp.shape == (48, 64) and (bmp >> 24).max() == 255
    # automatic pass size: about 8 Mi paths per pass, at most 64 samples (64 x 48 pixels: 64)
    log.clear()
    rt.Start(samples_per_pass=0, max_samples=100)
    assert len(log) == 4 and " 64.00/px " in log[2][0] and " 100.00/px " in log[3][0]
    # background thread + Pause / Resume / Stop handshake
    t = threading.Thread(target=lambda: rt.Start(samples_per_pass=1))
    t.start()
    time.sleep(0.3)
    assert rt.IsRunning
    rt.Pause()
    assert rt.IsPaused
    time.sleep(0.2)
    n1 = sum(rt.GetSampleSet(1, 1)[1:])
    time.sleep(0.2)
    assert sum(rt.GetSampleSet(1, 1)[1:]) == n1  # parked at a pass boundary
    rt.Resume()
    time.sleep(0.2)
    rt.Stop()
    t.join(timeout=20)
    assert not t.is_alive() and not rt.IsRunning and rt.IsStopping
    assert sum(rt.GetSampleSet(1, 1)[1:]) > n1
    rt.close()


def test_full_size_properties_1m_triangle_scene():
    """BASELINE C3 at full size (1M triangles, 2048x2048): size-independent properties of one wavefront."""
    sc = Scene.synthetic("soup", 1_000_000, 0xC3, 0.01)
    sc.override(width=2048, height=2048, recursion=4)
    ctx = Context(0, RTC_F32)
    ctx.load(sc, seed=1)
    ctx.set_option(RTC_OPT_COUNTERS, 1)
    ctx.render(0, 1)
    rgb, s, m = ctx.read_accum()
    assert np.all(s + m == 1) and np.isfinite(rgb).all() and (rgb >= 0).all()
    st = ctx.stats()
    assert st.paths == 2048 * 2048 and 2048 * 2048 <= st.rays <= 5 * 2048 * 2048
    # traversal work per ray stays at the tree's scale: one degenerate (non-finite) bounce ray that slipped through would
    # walk all 2.3e5 nodes, and a few hundred of them show up here (measured: 29.6 nodes, 5.4 primitives per ray)
    assert st.nodes_visited / st.rays < 34 and st.prims_tested / st.rays < 8
    ctx.set_option(RTC_OPT_COUNTERS, 0)
    assert 0.05 < m.mean() < 0.9  # the soup covers part of the frame
    ctx.render(1, 1)
    rgb2, s2, m2 = ctx.read_accum()
    ctx.clear_accum()
    ctx.render(0, 2)
    rgb3, s3, m3 = ctx.read_accum()
    assert np.array_equal(rgb2, rgb3) and np.array_equal(s2, s3) and np.array_equal(m2, m3)  # idempotent + additive
    ctx.close()


def test_odd_image_sizes_empty_scene_and_non_finite_rays():
    """Ragged sizes (1x1, a width that is no multiple of the warp, more bands than rows), the no-primitive scene
    (rejected with an error text, like the reference's missing accelerator assertion, Scene.cs:116) and rays with
    non-finite components (f32 mode: reported as misses, never traversed)."""
    for w, h in ((1, 1), (37, 5), (5, 67)):
        sc = cornell(w, h, 4)
        ora = O.OracleScene(sc, seed=9)
        ctx = Context(0, RTC_F64)
        ctx.set_option(N.RTC_OPT_MAX_PATHS, 1024)
        ctx.load(sc, seed=9)
        img = ctx.render_samples(2)
        ref = ora.render_samples(2)
        assert img.shape == (h, w, 3)
        assert np.isclose(img, ref, rtol=1e-9, atol=1e-12).all(axis=2).mean() >= (0.99 if w * h > 100 else 1.0)
        ctx.render(0, 3)
        rgb, s, m = ctx.read_accum()
        assert np.all(s + m == 3)
        got = ctx.render_read(3, 2)
        assert np.all(got[1] + got[2] == 5)
        ctx.close()
    empty = Scene.from_string("size 8 8\ncamera 0 0 -5 0 0 0 0 1 0 40\n")
    ctx = Context(0, RTC_F32)
    ctx.upload_scene(empty)
    with pytest.raises(N.RtcError) as e:
        ctx.build_bvh()
    assert e.value.code == N.RTC_ERR_INVALID and "no primitives" in str(e.value)
    ctx.close()
    sc = cornell(8, 8, 4)
    ctx = Context(0, RTC_F32)
    ctx.load(sc, seed=1)
    r = np.zeros(6, RAY_DT)
    r["origin"] = [[0, 0, -1]] * 6
    r["dir"] = [[0, 0, 1], [np.nan, 0, 1], [0, np.inf, 0], [0, 0, 1], [0, 0, 1], [0, 0, 1]]
    r["origin"][3] = [np.nan, 0, 0]
    r["origin"][4] = [0, -np.inf, 0]
    h = ctx.trace_closest(r)
    assert h["prim"][0] >= 0 and h["prim"][5] == h["prim"][0]
    assert np.all(h["prim"][1:5] == -1)
    ctx.close()


def test_render_read_equals_render_then_read():
    """rtc_render_read (band read-back overlapped with rendering, several bands) == rtc_render + rtc_read_accum."""
    sc = cornell(96, 80, 6)
    for prec in (RTC_F32, RTC_F64):
        ctx = Context(0, prec)
        ctx.set_option(N.RTC_OPT_MAX_PATHS, 96 * 16)  # 5 bands of 16 rows
        ctx.load(sc, seed=5)
        ctx.render(0, 3)
        want = ctx.read_accum()
        ctx.clear_accum()
        got = ctx.render_read(0, 3)
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
        got2 = ctx.render_read(3, 0)  # nothing to render: a plain read
        for a, b in zip(got2, want):
            assert np.array_equal(a, b)
        ctx.close()


def test_repeated_staged_uploads_and_frames():
    """The end-to-end frame loop of bench.py: re-upload the pinned scene image (its shading half arrives on the copy stream
    behind the first trace launch), clear, render + read back -- five times, alternating two different scenes in the same
    context, each frame bit-identical to a fresh context's."""
    scenes = [cornell(64, 48, 6), die(64, 48, 3)]
    want, baked = [], []
    for sc in scenes:
        c = Context(0, RTC_F32)
        c.load(sc, seed=4)
        c.render(0, 3)
        want.append(c.read_accum())
        baked.append(c.bake())
        c.close()
    ctx = Context(0, RTC_F32)
    for i in range(5):
        k = i & 1
        ctx.upload_baked(baked[k])
        ctx.set_params(scenes[k].params(4))
        ctx.set_camera(scenes[k].camera())
        ctx.clear_accum()
        got = ctx.render_read(0, 3)
        assert all(np.array_equal(a, b) for a, b in zip(got, want[k])), (i, k)
    ctx.close()
    for b in baked:
        b.close()


def test_baked_scene_image_round_trip():
    """rtc_bake / rtc_upload_baked: the host-resident device image renders exactly like the original hand-over."""
    sc = cornell(64, 64, 6)
    for prec in (RTC_F32, RTC_F64):
        ctx = Context(0, prec)
        ctx.load(sc, seed=3)
        ctx.render(0, 2)
        a = ctx.read_accum()
        baked = ctx.bake()
        assert baked.nbytes > 22 * 48
        ctx2 = Context(0, prec)
        ctx2.upload_baked(baked)
        ctx2.set_params(sc.params(3))
        ctx2.set_camera(sc.camera())
        ctx2.render(0, 2)
        b = ctx2.read_accum()
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        other = Context(0, RTC_F64 if prec == RTC_F32 else RTC_F32)
        with pytest.raises(N.RtcError):
            other.upload_baked(baked)  # an image belongs to one arithmetic mode
        with pytest.raises(N.RtcError):
            ctx2.get_bvh()  # no host-side tree behind a baked image
        for c in (ctx, ctx2, other):
            c.close()
        baked.close()


def test_debug_raycaster_overlays():
    """DebugRaycaster's per-pixel queries (primitive ids, BVH box counts) against the oracle."""
    for make in (cornell, die):
        sc = make(72, 48)
        ora = O.OracleScene(sc, seed=1)
        want_p, want_b = ora.debug_raycast(0), ora.debug_raycast(1)
        assert (want_p >= 0).mean() > 0.2 and want_b.max() > 5
        for prec in (RTC_F64, RTC_F32):
            ctx = Context(0, prec)
            ctx.load(sc, seed=1)
            got_p, got_b = ctx.debug_raycast(0), ctx.debug_raycast(1)
            assert np.array_equal(got_b, want_b)  # box counts are evaluated in f64 in both modes
            if prec == RTC_F64:
                assert np.array_equal(got_p, want_p)
            else:
                assert (got_p != want_p).mean() < 0.01  # unjittered rays land exactly on shared edges: f32 ties
            ctx.close()


def test_selection_overlay_shows_the_selected_primitives_alone():
    """DebugRaycaster's Selection mode over primitives (DebugRaycaster.cs:140-192): per pixel the nearest of the selected
    primitives, whatever hides it in the full scene. Checked against the oracle's primitive overlay of a scene that holds the
    selection only (f64: pixel for pixel), and against the full overlay: all primitives selected = DisplayMode.Primitives."""
    hdr = "size 96 64\ncamera 0 -7 1  0 0 0  0 0 1  45\ntwosided true\n"
    prims = ["sphere 0 0 0 1\n", "sphere 0 2 0 1.5\n", "sphere 1.5 -1 0.5 0.4\n",
             "vertex -3 3 -1\nvertex 3 3 -1\nvertex 0 3 3\ntri 0 1 2\n", "sphere -1.5 1 0 0.8\n", "plane 0 0 1 1\n"]
    full = Scene.from_string(hdr + "".join(prims))
    selection = [1, 3, 4]  # partly hidden behind primitives 0 and 2 in the full scene
    sub = Scene.from_string(hdr + "".join(prims[i] for i in selection))
    want_sub = O.OracleScene(sub, seed=1).debug_raycast(0)
    want = np.where(want_sub >= 0, np.array(selection)[np.maximum(want_sub, 0)], -1)
    for prec in (RTC_F64, RTC_F32):
        ctx = Context(0, prec)
        ctx.load(full, seed=1)
        got = ctx.debug_raycast_selection(selection)
        if prec == RTC_F64:
            assert np.array_equal(got, want)
        else:
            assert (got != want).mean() < 0.01
        overlay = ctx.debug_raycast(0)
        assert set(np.unique(got)) <= set(selection) | {-1}
        shown = np.isin(overlay, selection)
        assert np.array_equal(got[shown], overlay[shown])          # what is visible of the selection stays where it is
        assert (got >= 0).sum() > shown.sum()                      # ... and the hidden parts appear
        assert np.array_equal(ctx.debug_raycast_selection(list(range(full.n_prims))), overlay)
        assert np.all(ctx.debug_raycast_selection([]) == -1)
        with pytest.raises(N.RtcError):
            ctx.debug_raycast_selection([99])
        ctx.render(0, 1)  # the context's own scene is untouched
        ctx.close()


def test_ui_read_out_runs_beside_the_render_loop():
    """GetBitmap / GetSampleSet (rtc_tonemap_argb, rtc_read_pixel) polled from a second thread while
    the first thread renders: every read-out sees whole accumulation passes only, the final image equals an unpolled render
    bit for bit, and the render loop keeps its pace (passes per second with the poller within 25 % of without)."""
    sc = cornell(512, 512, 6)
    spp = 8
    ctx = Context(0, RTC_F32)
    ctx.load(sc, seed=2)
    ctx.render(0, 1)
    ctx.sync()

    def run(poll, seconds=None, passes=None):
        ctx.clear_accum()
        ctx.sync()
        stop = threading.Event()
        seen = []

        def poller():
            while not stop.is_set():
                img = ctx.tonemap(1.0, (0, 0, 0), 0.0)
                rgb, s, m = ctx.read_pixel(256, 400)
                seen.append((s + m, int(img[400, 256])))
                time.sleep(0.002)

        th = threading.Thread(target=poller)
        if poll:
            th.start()
        t0 = time.time()
        n = 0
        while (passes is not None and n < passes) or (seconds is not None and time.time() - t0 < seconds):
            ctx.render(n * spp, spp)
            ctx.sync()
            n += 1
        dt = time.time() - t0
        stop.set()
        if poll:
            th.join()
        return n / dt, n, seen, ctx.read_accum()

    run(False, seconds=0.2)  # warm-up
    rate_plain, _, _, _ = run(False, seconds=0.6)
    rate_poll, n, seen, _ = run(True, seconds=0.6)
    assert len(seen) >= 5
    counts = [c for c, _ in seen]
    assert all(c % spp == 0 and 0 <= c <= n * spp for c in counts) and counts == sorted(counts)  # whole passes, monotone
    print("render loop: %.1f passes/s unpolled, %.1f passes/s with %d read-outs" % (rate_plain, rate_poll, len(seen)))
    assert rate_poll >= 0.75 * rate_plain
    _, _, _, want = run(False, passes=12)
    _, _, _, got = run(True, passes=12)
    assert all(np.array_equal(a, b) for a, b in zip(got, want))
    # the read-out equals the synchronous path
    rgb, s, m = ctx.read_pixel(256, 400)
    assert np.array_equal(np.array(rgb), got[0][400, 256]) and s == got[1][400, 256] and m == got[2][400, 256]
    ctx.close()


def test_two_wavefronts_in_flight_render_the_same_planes():
    """RTC_OPT_WAVES: a frame large enough to be split over two streams (alternate bands, half the path pool each) must give
    the planes of the one-stream render bit for bit -- bands are disjoint pixel rows and every pixel's samples are still added
    in ascending order -- for a band count that is odd, a height that is no multiple of the band, and through rtc_render_read."""
    from raytracercore_b200 import RTC_OPT_MAX_PATHS, RTC_OPT_WAVES
    sc = cornell(1024, 1004, 4)
    out = {}
    for waves in (1, 2):
        ctx = Context(0, RTC_F32)
        ctx.set_option(RTC_OPT_WAVES, waves)
        ctx.set_option(RTC_OPT_MAX_PATHS, 8 << 20)  # 4 Mi paths per wavefront: two uneven bands, two sample chunks each
        ctx.load(sc, seed=3)
        ctx.render(0, 9)
        a = ctx.read_accum()
        b = ctx.render_read(9, 9)
        st = ctx.stats()
        out[waves] = (a, b, st.paths, st.rays)
        ctx.close()
    for k in range(2):
        assert all(np.array_equal(x, y) for x, y in zip(out[1][k], out[2][k]))
    assert out[1][2:] == out[2][2:]
    assert np.all(out[2][1][1] + out[2][1][2] == 18)


@pytest.mark.gpu
def test_queue_reordering_between_bounces_renders_the_same_planes():
    """RTC_OPT_REORDER sorts the live queue by the Morton cell of the ray origins (mode 1) or by cell and direction octant
    (mode 2) before every bounce after the first. Only the order in which paths are traced and shaded changes: planes, path and
    ray counts must equal the unsorted render bit for bit, with one and with two wavefronts in flight."""
    from raytracercore_b200 import RTC_OPT_MAX_PATHS, RTC_OPT_REORDER, RTC_OPT_WAVES
    sc = Scene.synthetic("spheres", 20000, 0xC4, 0.0)
    sc.override(width=512, height=384, recursion=5)
    out = {}
    for mode, waves in ((0, 1), (1, 1), (2, 1), (1, 2)):
        ctx = Context(0, RTC_F32)
        ctx.set_option(RTC_OPT_WAVES, waves)
        ctx.set_option(RTC_OPT_REORDER, mode)
        if waves == 2:
            ctx.set_option(RTC_OPT_MAX_PATHS, 8 << 20)
        ctx.load(sc, seed=5)
        ctx.render(0, 48 if waves == 2 else 3)
        st = ctx.stats()
        out[(mode, waves)] = (ctx.read_accum(), st.paths, st.rays)
        ctx.close()
    for key in ((1, 1), (2, 1)):
        assert all(np.array_equal(x, y) for x, y in zip(out[(0, 1)][0], out[key][0])), key
        assert out[(0, 1)][1:] == out[key][1:]
    ref = Context(0, RTC_F32)
    ref.set_option(RTC_OPT_WAVES, 1)
    ref.load(sc, seed=5)
    ref.render(0, 48)
    assert all(np.array_equal(x, y) for x, y in zip(ref.read_accum(), out[(1, 2)][0]))
    ref.close()


@pytest.mark.gpu
def test_create_horizon_on_the_device_against_the_oracle():
    """Vec4D.CreateHorizon (Vec4D.cs:33-58), evaluated directly: the lobe sample around a pole at polar
    cosine z and azimuth theta. f64: the oracle's value to rounding; f32: to 1e-5 (approximate sine / cosine / reciprocal
    square root). Includes the pole along the z axis (the (1, 0, 0) fall-back of the cross product) and z = 1."""
    import math
    rng = np.random.default_rng(11)
    n = 4096
    pole = rng.normal(size=(n, 3))
    pole /= np.linalg.norm(pole, axis=1, keepdims=True)
    pole[:4] = [[0, 0, 1], [0, 0, -1], [1, 0, 0], [0, 1, 0]]
    z = rng.uniform(0, 1, n)
    z[4:8] = [0.0, 1.0, 0.5, 0.999999]
    theta = rng.uniform(0, 2 * math.pi, n)
    theta[:2] = [0.0, math.pi / 2]
    tuples = np.concatenate([pole, z[:, None], theta[:, None]], axis=1)
    want = np.array([O.create_horizon(pole[i], z[i], theta[i]) for i in range(n)])
    for prec, tol in ((RTC_F64, 1e-12), (RTC_F32, 1e-5)):
        ctx = Context(0, prec)
        got = ctx.create_horizon(tuples)
        assert np.abs(got - want).max() < tol, (prec, np.abs(got - want).max())
        assert np.abs(np.linalg.norm(got, axis=1) - 1).max() < 10 * tol
        ctx.close()
