"""Pins the oracle against the hand-derived known answers of tests/golden/kat.json (see make_kat.py) and checks its
internal consistency: the accelerated walk (collect leaves, sort, early-out) equals the reference's plain loop."""
import json
import math
import os

import numpy as np
import pytest

import oracle as O
from conftest import SCENES
from parity import random_rays
from raytracercore_b200 import RAY_DT, Scene

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))


def test_philox_known_answers():
    for v in KAT["philox"]:
        assert list(O.philox(v["ctr"], v["key"])) == v["out"]


def test_uniform_mapping():
    u0, u1 = O.uniforms(0, 0, 0, 0, 0)
    r = KAT["philox"][0]["out"]
    assert u0 == ((r[1] << 32 | r[0]) >> 11) * 2.0 ** -53
    assert u1 == ((r[3] << 32 | r[2]) >> 11) * 2.0 ** -53
    us = [O.uniforms(7, p, s, st, b) for p in range(3) for s in range(3) for st in range(2) for b in range(2)]
    flat = [x for u in us for x in u]
    assert all(0.0 <= x < 1.0 for x in flat) and len(set(flat)) == len(flat)


@pytest.mark.parametrize("entry", KAT["scenes"], ids=[s["name"][:28] for s in KAT["scenes"]])
def test_intersection_known_answers(entry):
    sc = Scene.from_string("size 8 8\ncamera 0 0 -5 0 0 0 0 1 0 40\n" + entry["text"])
    ora = O.OracleScene(sc)
    rays = np.zeros(len(entry["cases"]), RAY_DT)
    for i, c in enumerate(entry["cases"]):
        rays["origin"][i] = c["origin"]
        rays["dir"][i] = c["dir"]
    for mode in (0, 1):
        got = ora.trace_closest(rays, mode=mode)
        for i, c in enumerate(entry["cases"]):
            w = c["hit"]
            assert got["prim"][i] == w["prim"], (entry["name"], i, got[i])
            if w["prim"] < 0:
                continue
            assert got["inside"][i] == w["inside"]
            assert got["t"][i] == pytest.approx(w["t"], rel=1e-12, abs=1e-12)
            assert np.allclose(got["position"][i], w["position"], atol=1e-12)
            assert np.allclose(got["normal"][i], w["normal"], atol=1e-12)


def test_aabb_known_answers():
    for v in KAT["aabb"]:
        ok, n, f = O.aabb_intersect(v["bmin"], v["bmax"], v["origin"], v["dir"])
        assert ok == v["hit"], v
        if ok:
            assert n == v["near"] and f == v["far"]


def test_tonemap_known_answers():
    for v in KAT["tonemap"]:
        out = O.tonemap(np.array([[v["rgb"]]], float), np.array([[v["samples"]]], np.uint32), np.array([[v["misses"]]], np.uint32),
                        v["exposure"], v["back"], v["back_a"])
        assert int(out[0, 0]) == v["argb"], (hex(int(out[0, 0])), hex(v["argb"]))


def test_create_horizon_properties():
    # Vec4D.CreateHorizon (Vec4D.cs:52-58): unit result at polar cosine z around the pole; pole along +z uses the (1,0,0) fallback
    rng = np.random.default_rng(3)
    for _ in range(50):
        pole = rng.normal(size=3)
        pole /= np.linalg.norm(pole)
        z, th = rng.uniform(0, 1), rng.uniform(0, 2 * math.pi)
        v = O.create_horizon(pole, z, th)
        assert abs(np.linalg.norm(v) - 1) < 1e-12 and abs(v @ pole - z) < 1e-12
    v = O.create_horizon([0, 0, 1], 0.0, 0.0)
    assert np.allclose(v, [1, 0, 0], atol=1e-15)
    v = O.create_horizon([0, 0, 1], 0.0, math.pi / 2)
    assert np.allclose(v, [0, 1, 0], atol=1e-15)


def _skip_cases(ora, rays):
    """Secondary rays leaving the first hit, with that hit as skipHit (Scene.RayTrace(ray, prevHit))."""
    first = ora.trace_closest(rays)
    m = first["prim"] >= 0
    rng = np.random.default_rng(5)
    sec = np.zeros(m.sum(), RAY_DT)
    sec["origin"] = first["position"][m]
    d = rng.normal(size=(m.sum(), 3))
    sec["dir"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    return sec, first[m]


@pytest.mark.parametrize("name", ["cornell_bounce.scene", "die.scene"])
def test_accelerated_walk_equals_plain_loop(name):
    sc = Scene.from_file(os.path.join(SCENES, name))
    ora = O.OracleScene(sc)
    rays = random_rays(np.random.default_rng(11), 20000, -2.5, 2.5, RAY_DT)
    _, diff = ora.trace_closest(rays, check_both=True)
    assert diff == 0
    sec, skip = _skip_cases(ora, rays)
    hits, diff = ora.trace_closest(sec, skip, check_both=True)
    assert diff == 0
    # a ray leaving a flat primitive never re-hits it (self-hit removal, Util.cs:179-192)
    kinds = sc.arrays()["kind"]
    flat = kinds[skip["prim"]] != 1
    assert not np.any(hits["prim"][flat] == skip["prim"][flat])


def test_self_hit_rule_lets_a_sphere_be_hit_again_from_inside():
    sc = Scene.from_string("size 8 8\ncamera 0 0 -5 0 0 0 0 1 0 40\ntwosided true\nsphere 0 0 0 1\n")
    ora = O.OracleScene(sc)
    r0 = np.zeros(1, RAY_DT)
    r0["origin"][0] = [0, 0, -5]
    r0["dir"][0] = [0, 0, 1]
    h0 = ora.trace_closest(r0)
    r1 = np.zeros(1, RAY_DT)
    r1["origin"][0] = h0["position"][0]
    r1["dir"][0] = [0, 0, 1]  # transmitted straight through
    h1 = ora.trace_closest(r1, h0)
    assert h1["prim"][0] == 0 and h1["inside"][0] == 1 and h1["t"][0] == pytest.approx(2.0)
    r1["dir"][0] = [0, 0, -1]  # reflected back: the sphere must not be hit again
    assert ora.trace_closest(r1, h0)["prim"][0] == -1


def test_render_accumulates_like_sample_sets():
    sc = Scene.from_file(os.path.join(SCENES, "cornell_bounce.scene"))
    sc.override(width=24, height=16, recursion=4)
    ora = O.OracleScene(sc, seed=3)
    rgb, s, m, rays = ora.render(0, 3)
    assert np.all(s + m == 3) and rays >= 24 * 16 * 3
    per = [ora.render_samples(k) for k in range(3)]
    acc = np.zeros_like(rgb)
    cnt = np.zeros_like(s)
    for img in per:
        miss = np.all(img == -1, axis=2)
        acc[~miss] += img[~miss]
        cnt += (~miss).astype(np.uint32)
    assert np.array_equal(cnt, s) and np.array_equal(acc, rgb)
    # threads do not change the result (Philox is keyed by pixel and sample)
    rgb1, s1, m1, _ = ora.render(0, 3, threads=1)
    assert np.array_equal(rgb1, rgb) and np.array_equal(s1, s)
