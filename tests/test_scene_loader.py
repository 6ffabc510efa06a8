"""SceneLoader mirror (raytracercore_b200/host/scene_loader.cpp) against the hand-derived tables of SURVEY.md appendix A
and the command coverage / error convention of the reference's SceneLoader.cs."""
import math
import os

import numpy as np
import pytest

from conftest import SCENES
from raytracercore_b200 import LoaderException, Scene
from raytracercore_b200 import _native as N

T, S, P = N.RTC_KIND_TRIANGLE, N.RTC_KIND_SPHERE, N.RTC_KIND_PLANE
MIR, TWO, INV, XF = N.RTC_FLAG_MIRROR, N.RTC_FLAG_TWOSIDED, N.RTC_FLAG_INVERT, N.RTC_FLAG_TRANSFORMED


def mat(a, i):
    m = a["material"][i]
    return dict(emission=tuple(m[0:3]), diffuse=tuple(m[3:6]), specular=tuple(m[6:9]), refraction=tuple(m[9:12]), ior=m[12], shininess=m[13])


def test_bounce_scene_matches_appendix_a1():
    sc = Scene.from_file(os.path.join(SCENES, "cornell_bounce.scene"))
    g = sc.globals()
    assert (g.width, g.height, g.recursion, g.n_cameras, g.n_prims, g.debug_geom) == (700, 700, 10, 8, 22, 0)
    assert tuple(g.background) == (0, 0, 0) and g.background_alpha == 0 and tuple(g.ambient) == (0, 0, 0)
    a = sc.arrays()
    assert list(a["kind"]) == [T] * 13 + [S] + [T] * 6 + [S, S]
    # ids 0-4 light box: two-sided mirrored triangles, emission 5
    for i in range(5):
        assert a["flags"][i] == MIR | TWO and mat(a, i)["emission"] == (5, 5, 5) and mat(a, i)["shininess"] == 100
    # ids 5-10 room: single-sided, inverted
    room = {5: ((.1, 1, .1), (.1, .35, .1), 250), 6: ((1, .1, .1), (.35, .1, .1), 250), 7: ((.9,) * 3, (.2,) * 3, 250),
            8: ((.9,) * 3, (.2,) * 3, 250), 9: ((.9,) * 3, (.2,) * 3, 250), 10: ((.4, .4, .9), (.4, .4, .9), 1000)}
    for i, (d, s, sh) in room.items():
        assert a["flags"][i] == MIR | INV
        assert mat(a, i)["diffuse"] == d and mat(a, i)["specular"] == s and mat(a, i)["shininess"] == sh and mat(a, i)["ior"] == 0
    # face normals of the room: -y, +y, +x, -x, -z, +z (Cube.cs:99-112 emits +x,-x,+y,-y,+z,-z; instances come one per line)
    normals = a["geom"][5:11, 9:12]
    assert np.allclose(np.abs(normals), [[0, 1, 0], [0, 1, 0], [1, 0, 0], [1, 0, 0], [0, 0, 1], [0, 0, 1]])
    centres = a["geom"][5:11, 0:3] + 0.5 * (a["geom"][5:11, 3:6] + a["geom"][5:11, 6:9])
    assert np.allclose(centres, [[0, -2, -1], [0, 2, -1], [2, 0, -1], [-2, 0, -1], [0, 0, -2], [0, 0, 0]])
    # id 13 plain sphere, id 20 transformed glass lens, id 21 mirror ball
    assert a["flags"][13] == TWO and np.allclose(a["geom"][13, :5], [-1, -1.25, -1, .5, .25])
    assert a["flags"][20] == TWO | XF and a["xform"][20] == 0 and mat(a, 20)["ior"] == 1.52 and mat(a, 20)["shininess"] == 1e5
    assert mat(a, 20)["refraction"] == (.9, .9, .9) and mat(a, 20)["specular"] == (.9, .9, .9) and mat(a, 20)["diffuse"] == (0, 0, 0)
    assert a["flags"][21] == TWO and mat(a, 21)["shininess"] == 1e6 and mat(a, 21)["ior"] == 0 and mat(a, 21)["refraction"] == (0, 0, 0)
    # lens matrices: MatrixToObject = translate(1,-1.25,-.5) rotate(-z,45) scale(.25,1,1); MatrixToWorld its inverse
    x = a["xforms"][0]
    to_world, to_object, to_normal = x[0:16].reshape(4, 4), x[16:32].reshape(4, 4), x[32:48].reshape(4, 4)
    assert np.allclose(to_world @ to_object, np.eye(4), atol=1e-14)
    c, s_ = math.cos(-math.pi / 4), math.sin(-math.pi / 4)
    rot = np.array([[c, -s_, 0], [s_, c, 0], [0, 0, 1]])  # rotation about +z by -45 deg == about -z by +45 deg
    assert np.allclose(to_object[:3, :3], rot @ np.diag([.25, 1, 1]), atol=1e-14) and np.allclose(to_object[:3, 3], [1, -1.25, -.5])
    assert np.allclose(to_normal[:3, :3], to_world[:3, :3].T)
    # ids 14-19: cube rotated 45 deg about z and moved to x = 1.25: vertices baked (Triangle.cs:68-74)
    cube_c = a["geom"][14:20, 0:3] + 0.5 * (a["geom"][14:20, 3:6] + a["geom"][14:20, 6:9])
    assert np.allclose(cube_c.mean(axis=0), [1.25, 0, -.375], atol=1e-12)
    # camera 0 (bounce.txt:13)
    cam = sc.camera(0, 512, 512)
    assert cam.kind == 0 and tuple(cam.position) == (2.8, -2.8, -1) and cam.dof_amount == 0 and cam.image_plane == 0
    assert cam.focal_length == pytest.approx(math.hypot(2.8, 2.8)) and cam.tan_fov_y2 == pytest.approx(-1.0) and cam.tan_fov_x2 == pytest.approx(1.0)
    assert np.allclose(cam.look, [-math.sqrt(.5), math.sqrt(.5), 0]) and cam.w2 == 256


def test_die_scene_matches_appendix_a2():
    sc = Scene.from_file(os.path.join(SCENES, "die.scene"))
    g = sc.globals()
    assert (g.width, g.height, g.recursion, g.n_cameras, g.n_prims) == (1280, 960, 3, 3, 29)
    a = sc.arrays()
    assert list(a["kind"]) == [S, S] + [T] * 6 + [S] * 21
    assert mat(a, 0)["emission"] == (10, 8, 8) and mat(a, 1)["emission"] == (8, 8, 10) and mat(a, 0)["diffuse"] == (0, 0, 0)
    faces = {2: (.9, 0, .9), 3: (.5, 1, .5), 4: (1, .5, .5), 5: (0, .7, .7), 6: (.7, .7, 0), 7: (.5, .5, 1)}
    for i, d in faces.items():
        assert mat(a, i)["diffuse"] == d and a["flags"][i] == MIR | TWO
    assert all(mat(a, i)["diffuse"] == (.9, .9, .9) and mat(a, i)["specular"] == (.5, .5, .5) and a["geom"][i, 3] == .15 for i in range(8, 29))
    assert np.all(a["material"][:, 13] == 100) and np.all(a["material"][:, 12] == 0) and np.all(a["xform"] == -1)
    cam = sc.camera(0, 1920, 1080)
    assert (cam.image_plane, cam.dof_amount, cam.focal_length) == (.1, 1000, 3) and tuple(cam.position) == (-2, -2, 2)
    assert cam.tan_fov_y2 == pytest.approx(-math.tan(math.radians(30))) and cam.tan_fov_x2 == pytest.approx(math.tan(math.radians(30)) * 1920 / 1080)


@pytest.mark.parametrize("mine,ref", [("cornell_bounce.scene", "bounce.txt"), ("die.scene", "die.txt")])
def test_reauthored_scene_equals_reference_file(mine, ref):
    path = os.path.join("/root/reference/Scenes", ref)
    if not os.path.exists(path):
        pytest.skip("reference checkout not present")
    a, b = Scene.from_file(os.path.join(SCENES, mine)), Scene.from_file(path)
    A, B = a.arrays(), b.arrays()
    for k in A:
        assert np.array_equal(A[k], B[k]), k
    ga, gb = a.globals(), b.globals()
    assert bytes(ga) == bytes(gb)
    for i in range(ga.n_cameras):
        assert bytes(a.camera(i)) == bytes(b.camera(i))


def test_command_coverage():
    text = """
    # comment only line
    SIZE 64, 32        # commands are case-insensitive, commas optional
    background .1 .2 .3 .5
    ambient miss
    bounce 7
    debug geom
    dof .5 12 at 0 0 3
    orthographic 0 0 -5, 0 0 0, 0 1 0, 4
    dof .25 6 camera
    frustum 1 2 3  0 0 0  0 0 1  60
    twosided yes
    invert 1
    emission 1 2 3
    diffuse .1 .2 .3
    specular .4 .5 .6
    shininess 2 10
    refraction .7 .8 .9 1.33
    maxverts 4
    vertex 0 0 0
    vertex 1 0 0
    vertex 0 1 0
    tri 0 1 2
    refraction off
    twosided false
    invert n
    vertexnormal 0 0 0  0 0 1
    vertexnormal 1 0 0  0 0 2
    vertexnormal 0 1 0  0 1 1
    trinormal 0 1 2
    plane 3 0 0 2
    cube 0 0 0 2 2 2 only +x -z
    cube 0 0 0 2 2 2 not x y z -x
    pushtransform
    translate 1 0 0
    scale 2 2 2
    rotate 0 0 1 90
    sphere 0 0 0 1
    poptransform
    sphere 0 0 0 1
    output ignored.png
    """
    sc = Scene.from_string(text)
    g = sc.globals()
    assert (g.width, g.height, g.recursion, g.debug_geom, g.n_cameras) == (64, 32, 7, 1, 2)
    assert tuple(g.ambient) == (-1, -1, -1) and tuple(g.background) == (.1, .2, .3) and g.background_alpha == .5
    c0, c1 = sc.camera(0), sc.camera(1)
    assert c0.kind == 1 and c0.focal_length == pytest.approx(math.sqrt(0 + 0 + 64)) and (c0.image_plane, c0.dof_amount) == (.5, 12)
    assert c0.h_mult == pytest.approx(4 / 32) and c0.v_mult == pytest.approx(-(1 / 16) * (32 / 64) * 4)
    assert c1.kind == 0 and c1.focal_length == pytest.approx(math.sqrt(14)) and (c1.image_plane, c1.dof_amount) == (.25, 6)
    a = sc.arrays()
    assert list(a["kind"]) == [T, T, P, T, T, T, T, S, S]
    assert a["flags"][0] == TWO | INV and mat(a, 0)["shininess"] == 1024 and mat(a, 0)["ior"] == 1.33 and mat(a, 0)["refraction"] == (.7, .8, .9)
    # `refraction off` stops applying refraction; later primitives keep the default (black, ior 0)
    assert a["flags"][1] == N.RTC_FLAG_VNORMALS and mat(a, 1)["ior"] == 0 and mat(a, 1)["refraction"] == (0, 0, 0)
    assert np.allclose(a["xforms"][a["xform"][1], 0:9], [0, 0, 1, 0, 0, 1, 0, math.sqrt(.5), math.sqrt(.5)])  # Vertex normalises
    assert np.allclose(a["geom"][1, 9:12], 0)  # face Normal is never computed for trinormal (Triangle.cs:59)
    assert np.allclose(a["geom"][2, :4], [0, 0, 1, 3])  # plane normal normalised
    assert a["flags"][3] == MIR and a["flags"][4] == MIR  # cube only +x -z  -> 2 faces, emitted +x then -z
    assert np.allclose(a["geom"][3, 9:12], [1, 0, 0]) and np.allclose(a["geom"][4, 9:12], [0, 0, -1])
    assert np.allclose(a["geom"][5, 9:12], [0, -1, 0]) and np.allclose(a["geom"][6, 9:12], [0, 0, -1])  # not x y z -x -> -y, -z
    assert a["flags"][7] == XF and a["xform"][7] >= 0 and a["flags"][8] == 0 and a["xform"][8] == -1
    to_object = a["xforms"][a["xform"][7], 16:32].reshape(4, 4)
    assert np.allclose(to_object @ [1, 0, 0, 1], [1, 2, 0, 1], atol=1e-12)  # translate . scale . rotate applied to (1,0,0)


def test_bare_cube_creates_nothing_and_instance_draws_faces():
    sc = Scene.from_string("cube 0 0 0 2 2 2\n")
    assert sc.n_prims == 0
    sc = Scene.from_string("cube 0 0 0 2 2 2\ninstance +x -x\ninstance all\n")
    assert sc.n_prims == 8


@pytest.mark.parametrize("text,cmd,line", [
    ("size 10\n", "size", 1),                                  # missing parameter
    ("\n\nsphere 0 0 zero 1\n", "sphere", 3),                  # FormatException
    ("ambient purple\n", "ambient", 1),
    ("vertex 0 0 0\ntri 0 1 2\n", "tri", 2),                   # index out of range
    ("instance +x\n", "instance", 1),                          # no object yet
    ("poptransform\npoptransform\n", "poptransform", 2),       # stack underflow
    ("cube 0 0 0 1 1 1 only +w\n", "cube", 1),
    ("size 99999999999 1\n", "size", 1),                       # OverflowException
    ("shininess inf\n", "shininess", 1),                       # spellings strtod takes but double.Parse does not
    ("shininess nan(1)\n", "shininess", 1),
    ("shininess 0x1p3\n", "shininess", 1),
    ("shininess 1e\n", "shininess", 1),
    ("shininess .\n", "shininess", 1),
    ("size 0x10 16\n", "size", 1),
    ("size 1.0 16\n", "size", 1),
])
def test_loader_exception_names_command_and_line(text, cmd, line):
    with pytest.raises(LoaderException) as e:
        Scene.from_string(text)
    assert "command %s on line %d" % (cmd, line) in str(e.value)


def test_number_grammar_is_dotnets():
    """double.Parse(InvariantCulture): Infinity / NaN symbols, group separators, exponents, a leading '+'."""
    sc = Scene.from_string("shininess Infinity\nrefraction 1 1 1 +1.5e0\ndiffuse .5 5. 1E+1\nsphere 0 0 0 1\n")
    m = sc.arrays()["material"][0]
    assert np.isinf(m[13]) and m[12] == 1.5 and list(m[3:6]) == [0.5, 5.0, 10.0]
    sc = Scene.from_string("shininess -infinity\nemission NaN 0 0\nsphere 0 0 0 1\n")
    m = sc.arrays()["material"][0]
    assert m[13] == -np.inf and np.isnan(m[0])


def test_line_format_errors_and_missing_file():
    with pytest.raises(LoaderException):
        Scene.from_string("sphere 0,0 0 1\n")  # a comma must be followed by whitespace (lineRegex, SceneLoader.cs:38)
    assert Scene.from_file("/nonexistent/scene.txt") is None  # the reference returns null (SceneLoader.cs:430-439)
    assert Scene.from_string("unknowncommand 1 2 3\n").n_prims == 0  # logged and skipped (:367-369)


def test_camera_init_render_basis():
    # Camera.InitRender (Camera.cs:54-63): look, side, up orthonormal; up overwritten; side negated last
    sc = Scene.from_string("size 100 50\ncamera 1 2 3  4 5 6  0 0 1  90\n")
    c = sc.camera(0)
    look, side, up = np.array(c.look), np.array(c.side), np.array(c.up)
    assert np.allclose(look, np.ones(3) / math.sqrt(3))
    for v in (look, side, up):
        assert abs(np.linalg.norm(v) - 1) < 1e-14
    assert abs(look @ side) < 1e-14 and abs(look @ up) < 1e-14 and abs(side @ up) < 1e-14
    assert np.allclose(np.cross(look, -np.array([0, 0, 1.0])) / np.linalg.norm(np.cross(look, [0, 0, -1.0])), -side)
    assert c.w2 == 50 and c.h2 == 25 and c.tan_fov_x2 == pytest.approx(2.0) and c.tan_fov_y2 == pytest.approx(-1.0)


def test_synthetic_scenes_are_deterministic():
    a = Scene.synthetic("soup", 2000, 0xC3, 0.01)
    b = Scene.synthetic("soup", 2000, 0xC3, 0.01)
    A, B = a.arrays(), b.arrays()
    assert all(np.array_equal(A[k], B[k]) for k in A)
    assert a.n_prims == 2000 and np.all(A["kind"] == T) and np.all(A["flags"] == TWO)
    g = A["geom"]
    assert np.all(np.abs(g[:, 0:3]) <= 1.01) and np.all(np.abs(g[:, 3:9]) <= 0.02)
    emissive = (A["material"][:, 0] == 8).mean()
    assert 0.002 < emissive < 0.03
    s = Scene.synthetic("spheres", 3000, 0xC4)
    m = s.arrays()["material"]
    assert np.all(m[0::3, 13] == 1e6) and np.all(m[1::3, 12] == 1.52) and np.all(m[2::3, 3] > 0) and np.all(m[0::200, 0] == 8)
    assert (s.width, s.height, s.recursion) == (1920, 1080, 8)
