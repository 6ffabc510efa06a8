"""The host BVH build that replaces BVH.Construct: reference-shaped tree, reference leaf boxes."""
import os

import numpy as np
import pytest

from conftest import SCENES
from raytracercore_b200 import Scene


def check_tree(sc):
    nodes, root = sc.bvh_array()
    n = sc.n_prims
    assert len(nodes) == 2 * n - 1
    seen = np.zeros(n, bool)
    depth = 0
    stack = [(root, 0)]
    visited = 0
    while stack:
        i, d = stack.pop()
        visited += 1
        nd = nodes[i]
        depth = max(depth, d)
        if nd["prim"] >= 0:
            assert nd["left"] == -1 and nd["right"] == -1 and not seen[nd["prim"]]
            seen[nd["prim"]] = True
            lo, hi = sc.primitive_bounds(int(nd["prim"]))
            assert np.array_equal(lo, nd["bmin"]) and np.array_equal(hi, nd["bmax"])  # AABB.CreateFromBounded
        else:
            l, r = nodes[nd["left"]], nodes[nd["right"]]
            assert np.array_equal(nd["bmin"], np.minimum(l["bmin"], r["bmin"]))  # AABB.Combine (AABB.cs:38-43)
            assert np.array_equal(nd["bmax"], np.maximum(l["bmax"], r["bmax"]))
            stack += [(int(nd["left"]), d + 1), (int(nd["right"]), d + 1)]
    assert seen.all() and visited == len(nodes)
    return depth


@pytest.mark.parametrize("name", ["cornell_bounce.scene", "die.scene"])
def test_tree_over_reference_scenes(name):
    check_tree(Scene.from_file(os.path.join(SCENES, name)))


def test_tree_over_synthetic_soup_is_shallow():
    sc = Scene.synthetic("soup", 30000, 1, 0.01)
    assert check_tree(sc) <= 40


def test_leaf_boxes_follow_the_reference_formulas():
    sc = Scene.from_string("twosided true\nvertex 0 0 0\nvertex 2 0 0\nvertex 1 1 0\ntri 0 1 2\ntri 0 1 2 mirrored\nsphere 1 2 3 .5\n"
                           "pushtransform\nscale 2 1 1\nsphere 0 0 0 1\npoptransform\nplane 1 0 0 1\nplane 1 1 1 1\n")
    lo, hi = sc.primitive_bounds(0)
    assert np.allclose(lo, [0, 0, 0]) and np.allclose(hi, [2, 1, 0])
    lo, hi = sc.primitive_bounds(1)  # the parallelogram includes v0 + e1 + e2 = (3,1,0) (Triangle.cs:240-241)
    assert np.allclose(lo, [0, 0, 0]) and np.allclose(hi, [3, 1, 0])
    lo, hi = sc.primitive_bounds(2)
    assert np.allclose(lo, [.5, 1.5, 2.5]) and np.allclose(hi, [1.5, 2.5, 3.5])
    lo, hi = sc.primitive_bounds(3)  # ellipsoid bounds (Sphere.cs:220-232)
    assert np.allclose(lo, [-2, -1, -1]) and np.allclose(hi, [2, 1, 1])
    lo, hi = sc.primitive_bounds(4)  # axis-aligned plane: flat in z, infinite in x,y (Plane.cs:68-74)
    assert lo[2] == hi[2] == 1 and np.all(np.isinf(lo[:2])) and np.all(np.isinf(hi[:2]))
    lo, hi = sc.primitive_bounds(5)
    assert np.all(np.isinf(lo)) and np.all(np.isinf(hi))


def test_planes_are_chained_above_the_root_in_id_order():
    sc = Scene.from_string("twosided true\nplane 1 0 0 1\nsphere 0 0 0 1\nsphere 3 0 0 1\nplane 2 0 1 0\n")
    nodes, root = sc.bvh_array()
    assert nodes[nodes[root]["left"]]["prim"] == 0          # first plane is the outermost left leaf
    inner = nodes[nodes[root]["right"]]
    assert nodes[inner["left"]]["prim"] == 3                # then the second plane
    rest = nodes[inner["right"]]
    assert rest["prim"] == -1 and np.all(np.isfinite(rest["bmin"])) and np.all(np.isfinite(rest["bmax"]))


def test_single_primitive_and_empty_scene():
    sc = Scene.from_string("sphere 0 0 0 1\n")
    nodes, root = sc.bvh_array()
    assert len(nodes) == 1 and root == 0 and nodes[0]["prim"] == 0
    nodes, root = Scene.from_string("size 4 4\n").bvh_array()
    assert len(nodes) == 0 and root == -1


def test_fast_bounds_equal_the_general_path_bit_for_bit():
    """DescPrimitiveBounds writes the triangle / plain-sphere cases out; every box must equal the one computed through a full
    Primitive (GetCenter / GetMaxCenterDistance), on the reference scenes, on synthetic soups and sphere fields, on parallelograms,
    on degenerate and on far-from-origin triangles, and non-finite input must take the general path."""
    import ctypes as C

    from raytracercore_b200 import _native as N

    def both(sc):
        d = sc.desc()
        out = np.zeros((2, d.n_prims, 6))
        for general in (0, 1):
            lo, hi = (C.c_double * 3)(), (C.c_double * 3)()
            for i in range(d.n_prims):
                assert N.lib.rtcs_desc_primitive_bounds(C.byref(d), i, general, lo, hi) == 0
                out[general, i, :3], out[general, i, 3:] = lo[:], hi[:]
        return out

    rng = np.random.default_rng(4)
    tris = "".join("vertex %s %s %s\n" % tuple(repr(float(c)) for c in v) for v in np.concatenate([
        rng.uniform(-1, 1, (60, 3)), rng.uniform(-1, 1, (30, 3)) * 1e6 + 1e9, rng.uniform(-1, 1, (30, 3)) * 1e-9, np.zeros((3, 3))]))
    text = "twosided true\n" + tris + "".join("tri %d %d %d%s\n" % (3 * k, 3 * k + 1, 3 * k + 2, " mirrored" if k % 3 == 0 else "") for k in range(41))
    text += "sphere 1 2 3 .5\npushtransform\nscale 2 1 1\nrotate 1 2 3 40\nsphere 0 0 0 1\npoptransform\nplane 1 0 0 1\nplane 1 1 1 1\n"
    scenes = [Scene.from_string(text), Scene.from_file(os.path.join(SCENES, "cornell_bounce.scene")), Scene.from_file(os.path.join(SCENES, "die.scene")),
              Scene.synthetic("soup", 3000, 7, 0.05), Scene.synthetic("spheres", 2000, 9, 0.0)]
    for sc in scenes:
        a = both(sc)
        assert a[0].tobytes() == a[1].tobytes()
    # non-finite geometry: same (non-finite) answer through the general path
    # (the loader rejects `inf`: the description is patched by hand)
    keep = Scene.synthetic("soup", 4, 1, 0.05)  # (the description points into the scene's own arrays)
    d = keep.desc()
    geom = np.ctypeslib.as_array(d.geom, shape=(4 * 12,))
    geom[0] = np.inf
    geom[12 + 3] = np.nan
    lo, hi, lo2, hi2 = ((C.c_double * 3)() for _ in range(4))
    for i in range(2):
        N.lib.rtcs_desc_primitive_bounds(C.byref(d), i, 0, lo, hi)
        N.lib.rtcs_desc_primitive_bounds(C.byref(d), i, 1, lo2, hi2)
        assert np.array_equal(np.array(lo[:] + hi[:]), np.array(lo2[:] + hi2[:]), equal_nan=True)
        assert not np.isfinite(np.array(lo[:] + hi[:])).all()
