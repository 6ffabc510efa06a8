"""Scene.Prepare on the device (rtc_prepare_device) against the host path it replaces:

* RTC_BUILDER_SAH builds the host builder's binned-SAH tree (host/bvh_builder.cpp) node for node and bit for bit;
* the device flatten (collapse to the quantised 8-wide tree, leaf slots, primitive / material records) produces the host
  flatten's image byte for byte -- for the SAH tree, for the clustering (PLOC) tree and for hand-made trees;
* the f64 parity mode goes through the same builder and keeps its bit-exact hits.

Scenes cover the edge cases the host tests cover: planes (kept out of the tree, chained above the root), a single primitive,
planes only, hundreds of coincident triangles (median fall-back with tied centroids: the ordering pass), transformed spheres
and vertex-normal triangles (transform rows), and the BASELINE-sized soup."""
import os
import time

import numpy as np
import pytest

from conftest import SCENES
from parity import random_rays
from raytracercore_b200 import RAY_DT, RTC_BUILDER_PLOC, RTC_BUILDER_SAH, RTC_F32, RTC_F64, Context, Scene
from raytracercore_b200 import _native as N
from test_gpu_intersection import MIXED

pytestmark = pytest.mark.gpu

NODE_DT = np.dtype([("bmin", "<f8", 3), ("bmax", "<f8", 3), ("left", "<i4"), ("right", "<i4"), ("prim", "<i4"), ("pad", "<i4")])


def tree_bytes(ctx):
    nodes, n, root = ctx.get_bvh()
    return bytes((N.BvhNode * n).from_buffer_copy(nodes)), n, root


def image(ctx):
    b = ctx.bake()
    try:
        return b.segments()
    finally:
        b.close()


def make_scene(name):
    hdr = "size 8 8\ncamera 0 0 -5 0 0 0 0 1 0 40\ntwosided true\n"
    if name == "mixed":
        return Scene.from_string(MIXED)
    if name == "cornell":
        return Scene.from_file(os.path.join(SCENES, "cornell_bounce.scene"))
    if name == "die":
        return Scene.from_file(os.path.join(SCENES, "die.scene"))
    if name == "soup":
        return Scene.synthetic("soup", 30000, 11, 0.03)
    if name == "spheres":
        return Scene.synthetic("spheres", 20000, 0xC4, 0.0)
    if name == "duplicates":  # 600 identical triangles + 1: coincident centroids all the way down
        tri = "vertex 0 0 0\nvertex 1 0 0\nvertex 0 1 0\nvertex 5 5 5\nvertex 6 5 5\nvertex 5 6 5\n" + "tri 0 1 2\n" * 600 + "tri 3 4 5\n"
        return Scene.from_string(hdr + tri)
    if name == "line":  # centroids equal on two axes, regular on the third: ties between bins and in the median split
        return Scene.from_string(hdr + "".join("sphere %d 0 0 0.4\n" % i for i in range(257)))
    if name == "single":
        return Scene.from_string(hdr + "sphere 0 0 0 1\n")
    if name == "two":
        return Scene.from_string(hdr + "sphere 0 0 0 1\nsphere 0 0 0 1\n")
    if name == "planes":
        return Scene.from_string(hdr + "plane 1 0 0 1\nplane 0 1 0 2\n")
    if name == "planes+1":
        return Scene.from_string(hdr + "plane 1 0 0 1\nsphere 0 0 0 1\nplane 0 1 0 2\nplane 0 0 1 3\n")
    if name == "vnormals":
        return Scene.from_string(hdr + "vertexnormal -1 -1 0  0 0 1\nvertexnormal 1 -1 0  1 0 1\nvertexnormal 0 1 0  0 1 1\ntrinormal 0 1 2\n"
                                 "sphere 3 0 0 1\npushtransform\ntranslate 0 2 0\nscale 1 2 3\nsphere 0 0 0 1\npoptransform\n")
    raise KeyError(name)


ALL = ["mixed", "cornell", "die", "soup", "spheres", "duplicates", "line", "single", "two", "planes", "planes+1", "vnormals"]


def assert_images_equal(a, b, what):
    assert a.keys() == b.keys()
    for k in a:
        assert len(a[k]) == len(b[k]), "%s: segment %s has %d vs %d bytes" % (what, k, len(a[k]), len(b[k]))
        if a[k] != b[k]:
            x, y = np.frombuffer(a[k], np.uint8), np.frombuffer(b[k], np.uint8)
            bad = np.flatnonzero(x != y)
            raise AssertionError("%s: segment %s differs in %d bytes, first at offset %d" % (what, k, len(bad), bad[0]))


@pytest.mark.parametrize("name", ALL)
def test_device_sah_tree_and_image_equal_the_host_path(name):
    sc = make_scene(name)
    host = Context(0, RTC_F32)
    host.upload_scene(sc)
    host.build_bvh()  # host binned SAH + host flatten
    dev = Context(0, RTC_F32)
    dev.upload_scene(sc)
    st = dev.prepare_device(RTC_BUILDER_SAH)
    ht, hn, hroot = tree_bytes(host)
    dt, dn, droot = tree_bytes(dev)
    assert (hn, hroot) == (dn, droot)
    if ht != dt:
        a, b = np.frombuffer(ht, NODE_DT), np.frombuffer(dt, NODE_DT)
        bad = [i for i in range(hn) if a[i].tobytes() != b[i].tobytes()]
        raise AssertionError("%s: %d of %d nodes differ, first %d: host %s device %s" % (name, len(bad), hn, bad[0], a[bad[0]], b[bad[0]]))
    assert_images_equal(image(host), image(dev), name)
    assert st.n_bounded <= sc.n_prims
    # and it traces: same hits through both contexts (same image, same kernel)
    rays = random_rays(np.random.default_rng(5), 4096, -2.5, 2.5, RAY_DT)
    a, b = host.trace_closest(rays), dev.trace_closest(rays)
    assert a.tobytes() == b.tobytes()
    host.close()
    dev.close()


@pytest.mark.parametrize("name", ["mixed", "soup", "duplicates", "planes+1"])
def test_device_flatten_of_the_clustering_tree_equals_the_host_flatten(name):
    """The PLOC build numbers its nodes in atomic order, so two builds differ in numbering: take the tree the device path
    made, hand it to the host flatten (rtc_upload_bvh) and compare the images."""
    sc = make_scene(name)
    dev = Context(0, RTC_F32)
    dev.upload_scene(sc)
    dev.prepare_device(RTC_BUILDER_PLOC, radius=8)
    nodes, n, root = dev.get_bvh()
    host = Context(0, RTC_F32)
    host.upload_scene(sc)
    host.upload_bvh(nodes, n, root)
    assert_images_equal(image(host), image(dev), name)
    host.close()
    dev.close()


def test_f64_mode_builds_on_the_device_and_keeps_its_hits():
    sc = make_scene("mixed")
    rays = random_rays(np.random.default_rng(7), 20000, -5, 5, RAY_DT)
    host = Context(0, RTC_F64)
    host.upload_scene(sc)
    host.build_bvh()
    dev = Context(0, RTC_F64)
    dev.upload_scene(sc)
    dev.prepare_device(RTC_BUILDER_SAH)
    assert tree_bytes(host) == tree_bytes(dev)
    assert host.trace_closest(rays).tobytes() == dev.trace_closest(rays).tobytes()
    host.close()
    dev.close()


def test_prepared_scene_renders_and_rebakes():
    """A scene prepared on the device renders the same frame as the host-prepared one, and its image -- read back by rtc_bake --
    uploads into a fresh context (FullRaytracer.Start with a cached Scene.Prepare)."""
    sc = make_scene("cornell")
    sc.override(width=96, height=96, recursion=6)
    frames = []
    for how in ("host", "device", "rebaked"):
        ctx = Context(0, RTC_F32)
        if how == "host":
            ctx.load(sc, seed=3)
        elif how == "device":
            ctx.load(sc, seed=3, device_prepare=RTC_BUILDER_SAH)
            baked = ctx.bake()
        else:
            ctx.upload_baked(baked)
            ctx.set_params(sc.params(3))
            ctx.set_camera(sc.camera())
        ctx.render(0, 4)
        frames.append(ctx.read_accum())
        ctx.close()
    baked.close()
    for f in frames[1:]:
        for x, y in zip(frames[0], f):
            assert np.array_equal(x, y)


def test_invalid_trees_are_refused_by_the_host_flatten_only():
    """rtc_prepare_device builds its own tree; the validation of foreign trees stays with rtc_upload_bvh. A second prepare on
    the same context replaces the first scene completely."""
    ctx = Context(0, RTC_F32)
    ctx.upload_scene(make_scene("soup"))
    ctx.prepare_device(RTC_BUILDER_SAH)
    a = image(ctx)
    small = make_scene("single")
    ctx.upload_scene(small)
    ctx.prepare_device(RTC_BUILDER_SAH)
    ref = Context(0, RTC_F32)
    ref.upload_scene(small)
    ref.build_bvh()
    assert_images_equal(image(ref), image(ctx), "single after soup")
    assert len(a["prims"]) == 30000 * 48
    ctx.close()
    ref.close()


def test_full_size_soup_prepares_on_the_device_like_on_the_host():
    """BASELINE C3: 1 M triangles. Same tree, same image; the phase times are printed for the record."""
    sc = Scene.synthetic("soup", 1000000, 0xC3, 0.01)
    t = time.time()
    host = Context(0, RTC_F32)
    host.upload_scene(sc)
    host.build_bvh()
    host.sync()
    t_host = time.time() - t
    t = time.time()
    dev = Context(0, RTC_F32)
    dev.upload_scene(sc)
    st = dev.prepare_device(RTC_BUILDER_SAH)
    dev.sync()
    t_dev = time.time() - t
    print("1M soup: host prepare %.3f s; device prepare %.3f s (boxes %.1f ms, build %.1f ms / %d levels, flatten %.1f ms, total %.1f ms; %d wide nodes, depth %d)"
          % (t_host, t_dev, st.boxes_ms, st.build_ms, st.build_levels, st.flatten_ms, st.total_ms, st.n_wide_nodes, st.wide_depth))
    assert tree_bytes(host) == tree_bytes(dev)
    assert_images_equal(image(host), image(dev), "soup1m")
    host.close()
    dev.close()


def test_prepare_device_error_behaviour():
    """Same conventions as the host path: an error code and a text, never a crash; the context stays usable."""
    from raytracercore_b200 import RtcError
    ctx = Context(0, RTC_F32)
    with pytest.raises(RtcError) as e:
        ctx.prepare_device(RTC_BUILDER_SAH)  # no scene yet
    assert e.value.code == N.RTC_ERR_STATE
    sc = make_scene("mixed")
    ctx.upload_scene(sc)
    with pytest.raises(RtcError) as e:
        ctx.prepare_device(7)  # unknown builder
    assert e.value.code == N.RTC_ERR_INVALID
    with pytest.raises(RtcError):
        ctx.bake()  # nothing prepared yet
    ctx.prepare_device(RTC_BUILDER_SAH)
    ctx.set_params(sc.params(1))
    ctx.set_camera(sc.camera())
    ctx.render(0, 1)
    rgb, s, m = ctx.read_accum()
    assert np.all(s + m == 1)
    # a new scene invalidates the tree (Scene.AddPrimitive -> ResetAccelerator, Scene.cs:58-63)
    ctx.upload_scene(make_scene("single"))
    with pytest.raises(RtcError):
        ctx.render(1, 1)
    with pytest.raises(RtcError):
        ctx.get_bvh()
    ctx.close()


def test_random_small_scenes_with_tied_centroids():
    """Many small scenes of spheres and triangles on a coarse lattice (equal centroids, equal bin boundaries, every size from 1
    to 40 and a few larger ones): the device tree and image must equal the host path's on each -- the median fall-back, the
    ordering pass and the two-primitive split all decide ties by primitive ID."""
    rng = np.random.default_rng(2024)
    hdr = "size 8 8\ncamera 0 0 -9 0 0 0 0 1 0 40\ntwosided true\n"
    sizes = list(range(1, 41)) + [63, 64, 65, 127, 128, 129, 257, 500]
    host, dev = Context(0, RTC_F32), Context(0, RTC_F32)
    for n in sizes:
        lines = []
        nv = 0
        for i in range(n):
            c = rng.integers(0, 4, 3)  # 64 lattice sites: many primitives share a centroid
            if rng.random() < 0.6:
                lines.append("sphere %d %d %d %s\n" % (c[0], c[1], c[2], ("0.25", "0.5")[int(rng.integers(0, 2))]))
            else:
                lines.append("vertex %d %d %d\nvertex %d.5 %d %d\nvertex %d %d.5 %d\ntri %d %d %d\n" % (c[0], c[1], c[2], c[0], c[1], c[2], c[0], c[1], c[2], nv, nv + 1, nv + 2))
                nv += 3
        if n % 7 == 0:
            lines.insert(n // 2, "plane 0 0 1 5\n")
        sc = Scene.from_string(hdr + "".join(lines))
        for ctx in (host, dev):
            ctx.upload_scene(sc)
        host.build_bvh()
        dev.prepare_device(RTC_BUILDER_SAH)
        assert tree_bytes(host) == tree_bytes(dev), n
        assert_images_equal(image(host), image(dev), "lattice scene of %d primitives" % n)
    host.close()
    dev.close()


def test_degenerate_and_badly_scaled_geometry():
    """Zero-area triangles (three equal vertices, collinear vertices), a zero-radius sphere, primitives a million units from the
    origin beside primitives a millionth of a unit across: Scene.Prepare on the device equals the host path on all of them, and
    the f64 mode still answers every ray like the oracle (bit-exact primitive and inside flag)."""
    import oracle as O
    from parity import check_hits
    hdr = "size 8 8\ncamera 0 0 -5 0 0 0 0 1 0 40\ntwosided true\n"
    text = hdr + (
        "vertex 0 0 0\nvertex 0 0 0\nvertex 0 0 0\ntri 0 1 2\n"                       # a point
        "vertex 1 1 1\nvertex 2 2 2\nvertex 3 3 3\ntri 3 4 5\n"                       # a line
        "sphere 0.5 0.5 0 0\n"                                                        # radius 0
        "sphere 1000000 0 0 2\nvertex 1000000 1 0\nvertex 1000001 1 0\nvertex 1000000 2 0\ntri 6 7 8\n"
        "sphere 0 0 0 0.000001\nvertex 0.000001 0 0\nvertex 0.000002 0 0\nvertex 0.000001 0.000001 0\ntri 9 10 11\n"
        "sphere -1 0 0 0.5\nsphere 1 0 0 0.5\nvertex -2 -2 1\nvertex 2 -2 1\nvertex 0 2 1\ntri 12 13 14\n")
    sc = Scene.from_string(text)
    host, dev = Context(0, RTC_F32), Context(0, RTC_F32)
    for c in (host, dev):
        c.upload_scene(sc)
    host.build_bvh()
    dev.prepare_device(RTC_BUILDER_SAH)
    assert tree_bytes(host) == tree_bytes(dev)
    assert_images_equal(image(host), image(dev), "degenerate geometry")
    rng = np.random.default_rng(8)
    rays = random_rays(rng, 8192, -3, 3, RAY_DT)
    far = random_rays(rng, 2048, -3, 3, RAY_DT)
    far["origin"] += [1000000, 0, 0]
    rays = np.concatenate([rays, far])
    want = O.OracleScene(sc).trace_closest(rays)
    f64 = Context(0, RTC_F64)
    f64.upload_scene(sc)
    f64.prepare_device(RTC_BUILDER_SAH)
    check_hits(f64.trace_closest(rays), want, 1e-5, exact=True)
    assert (want["prim"] >= 0).sum() > 1000
    # the f32 mode finds the same surfaces wherever the hit is resolvable at float precision (near the origin)
    near = rays[:8192]
    a, b = dev.trace_closest(near), want[:8192]
    agree = (a["prim"] == b["prim"]).mean()
    assert agree > 0.995, agree
    for c in (host, dev, f64):
        c.close()


def _walk_wide_tree(img, sc):
    """Decodes the quantised 8-wide tree of a baked f32 image and checks its structure against the scene: every bounded
    primitive is referenced by exactly one leaf slot, children are addressed by base + popcount, and the de-quantised box of
    every child contains -- with the builder's one-step padding -- the true boxes of everything below it. Returns (nodes
    visited, depth)."""
    qn = np.frombuffer(img["qnodes"], dtype=np.uint32).reshape(-1, 24)
    prim_id = np.frombuffer(img["prim_id"], dtype=np.int32)
    n = sc.n_prims
    boxes = np.array([np.concatenate(sc_bounds(sc, i)) for i in range(n)])
    seen = np.zeros(n, bool)
    visited, max_depth = 0, 0

    def node_boxes(row):
        p = row[0:3].view(np.float32).astype(np.float64)
        em = int(row[3])
        step = np.array([2.0 ** (((em >> (8 * a)) & 0xFF) - 127) for a in range(3)])
        q = row[8:20].view(np.uint8).reshape(6, 8).astype(np.float64)  # rows: lo x,y,z then hi x,y,z; column = slot
        lo = p[:, None] + q[0:3] * step[:, None]
        hi = p[:, None] + q[3:6] * step[:, None]
        return lo, hi, step, em >> 24, int(row[6]) & 0xFF

    def walk(i, depth):
        nonlocal visited, max_depth
        visited += 1
        max_depth = max(max_depth, depth)
        row = qn[i]
        lo, hi, step, imask, lmask = node_boxes(row)
        assert imask & lmask == 0
        child_base, prim_base = int(row[4]), int(row[5])
        total_lo, total_hi = np.full(3, np.inf), np.full(3, -np.inf)
        for s in range(8):
            bit = 1 << s
            if imask & bit:
                c = child_base + bin(imask & (bit - 1)).count("1")
                assert c > i  # breadth-first emission: children behind their parent
                clo, chi = walk(c, depth + 1)
            elif lmask & bit:
                slot = prim_base + bin(lmask & (bit - 1)).count("1")
                p = int(prim_id[slot])
                assert not seen[p]
                seen[p] = True
                clo, chi = boxes[p, :3], boxes[p, 3:]
            else:
                assert np.all(lo[:, s] > hi[:, s])  # empty slot: inverted box
                continue
            # one step of padding on either side (half a step covers the kernel's rounding of the folded 2^23 addend)
            assert np.all(lo[:, s] <= clo - 0.5 * step) and np.all(hi[:, s] >= chi + 0.5 * step), (i, s)
            total_lo, total_hi = np.minimum(total_lo, clo), np.maximum(total_hi, chi)
        return total_lo, total_hi

    import sys
    sys.setrecursionlimit(10000)
    walk(0, 1)
    bounded = np.isfinite(boxes).all(axis=1)
    assert np.array_equal(seen, bounded)
    return visited, max_depth


def sc_bounds(sc, i):
    import ctypes as C
    d = sc.desc()
    lo, hi = (C.c_double * 3)(), (C.c_double * 3)()
    assert N.lib.rtcs_desc_primitive_bounds(C.byref(d), i, 0, lo, hi) == 0
    return np.array(lo[:]), np.array(hi[:])


@pytest.mark.parametrize("name", ["mixed", "cornell", "soup", "spheres", "duplicates", "planes+1"])
def test_quantised_tree_is_conservative_and_complete(name):
    """The structure of the device-prepared image itself, independent of any kernel: see _walk_wide_tree."""
    sc = make_scene(name)
    ctx = Context(0, RTC_F32)
    ctx.upload_scene(sc)
    st = ctx.prepare_device(RTC_BUILDER_SAH)
    visited, depth = _walk_wide_tree(image(ctx), sc)
    assert visited == st.n_wide_nodes and depth == st.wide_depth
    ctx.close()
