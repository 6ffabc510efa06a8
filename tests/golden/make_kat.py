"""Writes tests/golden/kat.json — hand-derived known answers that pin the oracle (and, through it, the CUDA path).

The reference (Zaggy1024/RaytracerCore) ships no tests or golden vectors and cannot run in this image, so these
vectors are derived here in closed form from the reference's documented semantics (file:line in each entry), with
plain Python arithmetic that shares no code with oracle/ or with the CUDA kernels. Run: python tests/golden/make_kat.py
"""
import json
import math
import os

HERE = os.path.dirname(os.path.abspath(__file__))
kat = {"philox": [], "scenes": []}

# Philox4x32-10 known answers published with Random123 (kat_vectors), the generator this repo substitutes for the
# unseeded System.Random of Raytracer.cs:48.
kat["philox"] = [
    {"ctr": [0, 0, 0, 0], "key": [0, 0], "out": [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]},
    {"ctr": [0xffffffff] * 4, "key": [0xffffffff] * 2, "out": [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]},
    {"ctr": [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], "key": [0xa4093822, 0x299f31d0],
     "out": [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]},
]

S2 = math.sqrt(0.5)


def hit(prim, inside, t, pos, normal):
    return {"prim": prim, "inside": inside, "t": t, "position": pos, "normal": normal}


MISS = {"prim": -1}

# --- unit sphere at the origin, two-sided (Sphere.cs:50-155) --------------------------------------------------
kat["scenes"].append({
    "name": "unit sphere",
    "text": "twosided true\nsphere 0 0 0 1\n",
    "cases": [
        # from outside along +z: near root t = 5 - 1, outward normal at the near point
        {"origin": [0, 0, -5], "dir": [0, 0, 1], "hit": hit(0, 0, 4.0, [0, 0, -1], [0, 0, -1])},
        # from the centre: tNear = -1 < 0, so only the far hit is returned, inside = true, normal flipped inwards (:142,:148)
        {"origin": [0, 0, 0], "dir": [1, 0, 0], "hit": hit(0, 1, 1.0, [1, 0, 0], [-1, 0, 0])},
        # offset chord: x = 0.6 -> half chord 0.8
        {"origin": [0.6, 0, -3], "dir": [0, 0, 1], "hit": hit(0, 0, 3 - 0.8, [0.6, 0, -0.8], [0.6, 0, -0.8])},
        # pointing away: both roots negative -> no hit (:145)
        {"origin": [0, 0, -5], "dir": [0, 0, -1], "hit": MISS},
        # passes at distance 1.5 from the centre: negative discriminant -> NaN -> no hit
        {"origin": [1.5, 0, -5], "dir": [0, 0, 1], "hit": MISS},
    ]})

# --- single-sided sphere: inside hits are culled after `Invert` (Primitive.cs:60-64) -----------------------------
kat["scenes"].append({
    "name": "one-sided sphere from inside",
    "text": "twosided false\nsphere 0 0 0 1\n",
    "cases": [
        {"origin": [0, 0, 0], "dir": [1, 0, 0], "hit": MISS},
        {"origin": [0, 0, -5], "dir": [0, 0, 1], "hit": hit(0, 0, 4.0, [0, 0, -1], [0, 0, -1])},
    ]})
kat["scenes"].append({
    "name": "inverted one-sided sphere: only visible from inside, Inside flag flipped, normal NOT flipped (Hit.cs:39-42)",
    "text": "twosided false\ninvert true\nsphere 0 0 0 1\n",
    "cases": [
        # from outside: near hit becomes inside=true -> culled; far hit (raw inside=true) becomes inside=false -> accepted
        {"origin": [0, 0, -5], "dir": [0, 0, 1], "hit": hit(0, 0, 6.0, [0, 0, 1], [0, 0, -1])},
        {"origin": [0, 0, 0], "dir": [1, 0, 0], "hit": hit(0, 0, 1.0, [1, 0, 0], [-1, 0, 0])},
    ]})

# --- triangle and parallelogram (Triangle.cs:77-146): v0=(0,0,0) e1=(1,0,0) e2=(0,1,0), N=+z -----------------------
tri = "twosided true\nvertex 0 0 0\nvertex 1 0 0\nvertex 0 1 0\n"
kat["scenes"].append({
    "name": "triangle",
    "text": tri + "tri 0 1 2\n",
    "cases": [
        # from +z looking down: det = e1.(D x e2) = 1*( (0*0 - (-1)*1) ) ... inv > 0 -> outside, normal +z
        {"origin": [0.25, 0.25, 1], "dir": [0, 0, -1], "hit": hit(0, 0, 1.0, [0.25, 0.25, 0], [0, 0, 1])},
        # from -z looking up: back face -> inside = true, normal flipped (:221-223)
        {"origin": [0.25, 0.25, -2], "dir": [0, 0, 1], "hit": hit(0, 1, 2.0, [0.25, 0.25, 0], [0, 0, -1])},
        # u + v > 1 -> outside the triangle (:119)
        {"origin": [0.75, 0.75, 1], "dir": [0, 0, -1], "hit": MISS},
        # behind the origin: t < 0 (:120)
        {"origin": [0.25, 0.25, 1], "dir": [0, 0, 1], "hit": MISS},
    ]})
kat["scenes"].append({
    "name": "mirrored triangle = parallelogram (u <= 1 and v <= 1, Triangle.cs:117-118)",
    "text": tri + "tri 0 1 2 mirrored\n",
    "cases": [
        {"origin": [0.75, 0.75, 1], "dir": [0, 0, -1], "hit": hit(0, 0, 1.0, [0.75, 0.75, 0], [0, 0, 1])},
        {"origin": [1.25, 0.5, 1], "dir": [0, 0, -1], "hit": MISS},
    ]})

# --- plane z = 2 (Plane.cs:36-66): Distance is |P - O|, inside when the ray runs along the normal --------------------
kat["scenes"].append({
    "name": "plane",
    "text": "twosided true\nplane 2 0 0 1\n",
    "cases": [
        {"origin": [0, 0, 0], "dir": [0, 0, 1], "hit": hit(0, 1, 2.0, [0, 0, 2], [0, 0, -1])},
        {"origin": [0, 0, 5], "dir": [0, 0, -1], "hit": hit(0, 0, 3.0, [0, 0, 2], [0, 0, 1])},
        {"origin": [1, 1, 5], "dir": [S2, 0, -S2], "hit": hit(0, 0, 3 / S2, [4, 1, 2], [0, 0, 1])},
        {"origin": [0, 0, 0], "dir": [1, 0, 0], "hit": MISS},  # parallel, off the plane (:44-45)
    ]})

# --- closest of several, strict `<` (Scene.cs:85-86) -----------------------------------------------------------
kat["scenes"].append({
    "name": "two spheres in a row + a plane behind",
    "text": "twosided true\nsphere 0 0 0 1\nsphere 0 0 4 1\nplane 10 0 0 1\n",
    "cases": [
        {"origin": [0, 0, -5], "dir": [0, 0, 1], "hit": hit(0, 0, 4.0, [0, 0, -1], [0, 0, -1])},
        {"origin": [0, 0, 2], "dir": [0, 0, 1], "hit": hit(1, 0, 1.0, [0, 0, 3], [0, 0, -1])},
        {"origin": [3, 0, 0], "dir": [0, 0, 1], "hit": hit(2, 1, 10.0, [3, 0, 10], [0, 0, -1])},
        {"origin": [0, 0, 2], "dir": [0, 0, -1], "hit": hit(0, 0, 1.0, [0, 0, 1], [0, 0, 1])},
    ]})

# --- transformed sphere (Sphere.cs:29-37,58-75,100-139): translate(2,0,0) scale(2,1,1) of the unit sphere ------------
kat["scenes"].append({
    "name": "ellipsoid",
    "text": "twosided true\npushtransform\ntranslate 2 0 0\nscale 2 1 1\nsphere 0 0 0 1\npoptransform\n",
    "cases": [
        # along x: surface at x = 2 - 2 = 0 ; distance re-measured in world space (:136)
        {"origin": [-3, 0, 0], "dir": [1, 0, 0], "hit": hit(0, 0, 3.0, [0, 0, 0], [-1, 0, 0])},
        # along z through the centre: semi-axis 1 in z
        {"origin": [2, 0, -4], "dir": [0, 0, 1], "hit": hit(0, 0, 3.0, [2, 0, -1], [0, 0, -1])},
        {"origin": [2, 0, 0], "dir": [-1, 0, 0], "hit": hit(0, 1, 2.0, [0, 0, 0], [1, 0, 0])},
    ]})

# --- AABB.IntersectAVX (AABB.cs:107-142): (bmin, bmax, origin, dir) -> hit?, near, far ---------------------------
kat["aabb"] = [
    {"bmin": [-1, -1, -1], "bmax": [1, 1, 1], "origin": [0, 0, -5], "dir": [0, 0, 1], "hit": True, "near": 4.0, "far": 6.0},
    # zero direction components with the origin inside those slabs are replaced by (-inf, +inf) (:117-123)
    {"bmin": [-1, -1, -1], "bmax": [1, 1, 1], "origin": [2, 0, -5], "dir": [0, 0, 1], "hit": False},
    # flat box (zero thickness in z): near == far passes, miss only if near > far (:138)
    {"bmin": [-1, -1, 0], "bmax": [1, 1, 0], "origin": [0, 0, -2], "dir": [0, 0, 1], "hit": True, "near": 2.0, "far": 2.0},
    # box behind the ray: far < 0
    {"bmin": [-1, -1, -1], "bmax": [1, 1, 1], "origin": [0, 0, 5], "dir": [0, 0, 1], "hit": False},
    # origin inside: near < 0 <= far
    {"bmin": [-1, -1, -1], "bmax": [1, 1, 1], "origin": [0, 0, 0], "dir": [1, 0, 0], "hit": True, "near": -1.0, "far": 1.0},
    # negative direction swaps the slabs (:126-127)
    {"bmin": [1, 1, 1], "bmax": [2, 3, 4], "origin": [5, 2, 2], "dir": [-1, 0, 0], "hit": True, "near": 3.0, "far": 4.0},
]

# --- SampleSet.GetOutput (SampleSet.cs:61-113): (rgb sum, samples, misses, exposure, back rgb, back alpha) -> ARGB --
def code(r, g, b, a):
    c = lambda v: int(min(max(v if v == v else 0.0, 0.0), 1.0) * 255)
    return (c(a) << 24) | (c(r) << 16) | (c(g) << 8) | c(b)

g = 1 / 2.2
kat["tonemap"] = [
    {"rgb": [4.0, 2.0, 1.0], "samples": 4, "misses": 0, "exposure": 1.0, "back": [0, 0, 0], "back_a": 0.0,
     "argb": code(1.0 ** g, 0.5 ** g, 0.25 ** g, 1.0)},
    # no samples: background * exposure, no gamma (:57-58)
    {"rgb": [0, 0, 0], "samples": 0, "misses": 3, "exposure": 2.0, "back": [0.25, 0.5, 0.1], "back_a": 0.5,
     "argb": code(0.5, 1.0, 0.2, 0.5)},
    # half the samples missed, opaque-ish background: colour blends toward back by misses/total*backA, alpha 1+(backA-1)*m/total
    {"rgb": [1.0, 1.0, 1.0], "samples": 2, "misses": 2, "exposure": 1.0, "back": [0.0, 0.0, 1.0], "back_a": 1.0,
     "argb": code((0.5 + (0 - 0.5) * 0.5) ** g, (0.5 + (0 - 0.5) * 0.5) ** g, (0.5 + (1 - 0.5) * 0.5) ** g, 1.0)},
    {"rgb": [1.0, 1.0, 1.0], "samples": 2, "misses": 2, "exposure": 1.0, "back": [0.0, 0.0, 0.0], "back_a": 0.0,
     "argb": code(0.5 ** g, 0.5 ** g, 0.5 ** g, 0.5)},
    # over-exposed and negative channels clamp to [0,1]; (int)(x*255) truncates
    {"rgb": [10.0, -1.0, 0.999], "samples": 1, "misses": 0, "exposure": 1.0, "back": [0, 0, 0], "back_a": 0.0,
     "argb": code(1.0, 0.0, 0.999 ** g, 1.0)},
]

with open(os.path.join(HERE, "kat.json"), "w") as f:
    json.dump(kat, f, indent=1)
print("wrote kat.json:", sum(len(s["cases"]) for s in kat["scenes"]), "intersection cases")
