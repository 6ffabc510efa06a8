"""Multi-GPU correctness on hardware: R contexts on R GPUs of the box, scene replicated, rank r renders
its sample range of every frame, one NCCL collective per frame (rtc_reduce_accum). The reduced SampleSet planes must equal
the single-GPU frame of R x spp samples: counters bit-exact, f64 colour sums to rounding. Progressive frames must keep
accumulating (no clear between frames), with ncclReduce to a root and with the all-reduce variant. Skips on a 1-GPU box
(the driver's scaling run covers the same check inside bench.py)."""
import os
import threading

import numpy as np
import pytest

from conftest import SCENES
from raytracercore_b200 import RTC_F32, Context, Scene
from raytracercore_b200 import _native as N
from raytracercore_b200.partition import frame_samples, sample_range

pytestmark = pytest.mark.gpu


def _run_ranks(world, fn):
    errs = []

    def wrap(r):
        try:
            fn(r)
        except Exception as e:  # noqa: BLE001
            errs.append((r, e))

    ts = [threading.Thread(target=wrap, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not errs, errs
    assert not any(t.is_alive() for t in ts)


@pytest.mark.parametrize("root", [0, -1])
def test_reduced_frames_equal_the_single_gpu_render(root):
    n_dev = N.lib.rtc_device_count()
    if n_dev < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(n_dev, 4)
    sc = Scene.from_file(os.path.join(SCENES, "cornell_bounce.scene"))
    sc.override(width=96, height=64, recursion=6)
    spp, frames = 3, 3
    single = Context(0, RTC_F32)
    single.load(sc, seed=12)
    want = []
    for f in range(frames):
        first, count = frame_samples(f, world, spp)
        single.render(first, count)
        want.append(single.read_accum())
    single.close()
    uid = Context.comm_unique_id()
    ctxs = [Context(r, RTC_F32) for r in range(world)]
    got = [[None] * frames for _ in range(world)]

    def rank_main(r):
        c = ctxs[r]
        c.load(sc, seed=12)
        c.comm_init(world, r, uid)
        for f in range(frames):
            c.render(*sample_range(f, r, world, spp))  # no clear between frames: the job's total keeps accumulating
            c.reduce_accum(root)
            got[r][f] = c.read_accum()

    _run_ranks(world, rank_main)
    holders = range(world) if root < 0 else [root]
    for f in range(frames):
        w_rgb, w_s, w_m = want[f]
        for r in holders:
            rgb, s, m = got[r][f]
            assert np.array_equal(s, w_s) and np.array_equal(m, w_m), (root, f, r)
            assert np.all(s + m == (f + 1) * world * spp)
            assert np.allclose(rgb, w_rgb, rtol=1e-12, atol=1e-12), (root, f, r)
        if root >= 0:  # the other ranks have handed their contribution over
            for r in range(world):
                if r != root:
                    assert not got[r][f][1].any() and not got[r][f][2].any() and not got[r][f][0].any()
    for c in ctxs:
        c.close()


def test_scene_broadcast_over_nvlink_gives_every_rank_the_roots_scene():
    """rtc_bcast_scene: only rank 0 prepares and uploads the scene; the other ranks receive the device image with ncclBroadcast
    and, given the same camera and parameters, must render bit-identical planes -- for a scene that replaces an earlier one
    of a different size on the receivers as well, and whether the root prepared the scene on the host or on the device
    (rtc_prepare_device). A receiving context has no host-side description, but rtc_bake reads its device image back: byte for
    byte the root's."""
    n_dev = N.lib.rtc_device_count()
    if n_dev < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(n_dev, 4)
    scenes = [Scene.from_file(os.path.join(SCENES, "cornell_bounce.scene")), Scene.synthetic("soup", 30000, 5, 0.02),
              Scene.from_file(os.path.join(SCENES, "die.scene"))]
    for sc in scenes:
        sc.override(width=80, height=48, recursion=4)
    uid = Context.comm_unique_id()
    ctxs = [Context(r, RTC_F32) for r in range(world)]
    got = [[None] * len(scenes) for _ in range(world)]
    images = [[None] * len(scenes) for _ in range(world)]

    def rank_main(r):
        from raytracercore_b200 import RTC_BUILDER_SAH
        c = ctxs[r]
        c.comm_init(world, r, uid)
        for k, sc in enumerate(scenes):
            if r == 0:
                c.load(sc, seed=4, device_prepare=RTC_BUILDER_SAH if k == 1 else None)
            c.bcast_scene(0)
            if r != 0:
                c.set_params(sc.params(4))
                c.set_camera(sc.camera())
            b = c.bake()
            images[r][k] = b.segments()
            b.close()
            c.clear_accum()
            c.render(0, 2)
            got[r][k] = c.read_accum()

    _run_ranks(world, rank_main)
    for k in range(len(scenes)):
        for r in range(1, world):
            assert all(np.array_equal(a, b) for a, b in zip(got[r][k], got[0][k])), (k, r)
            assert images[r][k] == images[0][k], (k, r)
        assert got[0][k][1].any()
    for c in ctxs:
        c.close()
