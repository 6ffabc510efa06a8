"""The P/Invoke shim a maintainer of the reference would add (host_csharp/) cannot be compiled here (no .NET SDK), so its
contract with the C ABI is checked textually: every `rtc_*` function include/rtcore_b200.h declares has a DllImport with the
same number of parameters and compatible types, every struct has the header's fields in the header's order with the same
sizes (= the ctypes structs the tests drive the library with), and GpuFullRaytracer exposes the whole public surface of the
reference's FullRaytracer that MainWindow.cs / RayInspector.cs call."""
import ctypes as C
import os
import re

from raytracercore_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "rtcore_b200.h")).read()
NATIVE = open(os.path.join(ROOT, "host_csharp", "RtcoreNative.cs")).read()
SHIM = open(os.path.join(ROOT, "host_csharp", "GpuFullRaytracer.cs")).read()

CS_SIZES = {"int": 4, "uint": 4, "long": 8, "ulong": 8, "double": 8, "byte": 1, "float": 4}


def strip_comments(text):
    return re.sub(r"//[^\n]*", "", re.sub(r"/\*.*?\*/", "", text, flags=re.S))


def header_functions():
    out = {}
    for m in re.finditer(r"^\s*(?:const\s+)?(\w+\*?)\s+(rtc_\w+)\(([^)]*)\);", strip_comments(HEADER), flags=re.M):
        args = [a.strip() for a in m.group(3).split(",")] if m.group(3).strip() not in ("", "void") else []
        out[m.group(2)] = (m.group(1), args)
    return out


def cs_imports():
    out = {}
    for m in re.finditer(r"\[DllImport\(Lib\)\]\s*public static extern (\w+\*?)\s+(rtc_\w+)\(([^)]*)\);", NATIVE):
        args = [a.strip() for a in m.group(3).split(",")] if m.group(3).strip() else []
        out[m.group(2)] = (m.group(1), args)
    return out


def c_kind(decl):
    """Width class of a C parameter: 'ptr', 4 or 8."""
    d = decl.replace("const ", "").strip()
    if "*" in d or "[" in d:
        return "ptr"
    t = d.split()[0]
    return {"int": 4, "int32_t": 4, "uint32_t": 4, "int64_t": 8, "uint64_t": 8, "double": 8}[t]


def cs_kind(decl):
    d = decl.strip()
    if d.startswith("out ") or "*" in d or d.startswith("IntPtr"):
        return "ptr"
    return CS_SIZES[d.split()[0]]


def test_every_header_function_has_a_matching_dllimport():
    hdr, cs = header_functions(), cs_imports()
    assert len(hdr) >= 38, sorted(hdr)
    missing = sorted(set(hdr) - set(cs))
    assert not missing, "no DllImport for: %s" % missing
    extra = sorted(set(cs) - set(hdr))
    assert not extra, "DllImport without a header declaration: %s" % extra
    assert set(hdr) == {k for k in N.SIGNATURES if k.startswith("rtc_")}  # and ctypes binds the same set
    for name, (ret, args) in hdr.items():
        cret, cargs = cs[name]
        assert len(args) == len(cargs), (name, args, cargs)
        for a, b in zip(args, cargs):
            assert c_kind(a) == cs_kind(b), (name, a, b)
        want = "ptr" if "*" in ret else {"int": 4, "int64_t": 8, "void": 0}[ret]
        got = "ptr" if cret == "IntPtr" else {"int": 4, "long": 8, "void": 0}[cret]
        assert want == got, (name, ret, cret)


def header_structs():
    out = {}
    for m in re.finditer(r"typedef struct (\w+) \{(.*?)\} \1;", strip_comments(HEADER), flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            tm = re.match(r"(const\s+)?(\w+)\s*(\*?)\s*(.*)", decl)
            base, ptr, names = tm.group(2), tm.group(3), tm.group(4)
            for nm in names.split(","):
                nm = nm.strip()
                star = ptr or ("*" if nm.startswith("*") else "")
                nm = nm.lstrip("* ")
                am = re.match(r"(\w+)\[(\w+)\]", nm)
                count = 1
                if am:
                    nm, count = am.group(1), {"RTC_K_COUNT": 5}.get(am.group(2)) or int(am.group(2))
                fields.append((nm, "ptr" if star else base, count))
        out[m.group(1)] = fields
    return out


def cs_structs():
    out = {}
    for m in re.finditer(r"public (?:unsafe )?struct (\w+)", NATIVE):
        i = NATIVE.index("{", m.end())
        j = NATIVE.index("}", i)  # the binding's structs hold no nested braces
        fields = []
        for decl in NATIVE[i + 1:j].split(";"):
            decl = decl.strip()
            if not decl.startswith("public"):
                continue
            decl = decl[len("public"):].strip()
            if decl.startswith("fixed "):
                decl = decl[len("fixed "):]
            tm = re.match(r"(\w+\*?)\s+(.*)", decl)
            typ, names = tm.group(1), tm.group(2)
            for nm in names.split(","):
                nm = nm.strip()
                am = re.match(r"(\w+)\[(\d+)\]", nm)
                count = int(am.group(2)) if am else 1
                fields.append((am.group(1) if am else nm, "ptr" if typ.endswith("*") else typ, count))
        out[m.group(1)] = fields
    return out


C_SIZES = {"int32_t": 4, "uint32_t": 4, "int64_t": 8, "uint64_t": 8, "double": 8, "uint8_t": 1, "ptr": 8, "rtc_hit": None}
PAIRS = {"rtc_scene_desc": ("RtcSceneDesc", N.SceneDesc), "rtc_bvh_node": ("RtcBvhNode", N.BvhNode), "rtc_camera": ("RtcCamera", N.Camera),
         "rtc_params": ("RtcParams", N.Params), "rtc_ray": ("RtcRay", N.Ray), "rtc_hit": ("RtcHit", N.Hit),
         "rtc_debug_ray": ("RtcDebugRay", N.DebugRay), "rtc_stats": ("RtcStats", N.Stats),
         "rtc_prepare_stats": ("RtcPrepareStats", N.PrepareStats)}


def test_struct_layouts_match_the_header_and_the_ctypes_view():
    hs, cs = header_structs(), cs_structs()
    for cname, (csname, ctype) in PAIRS.items():
        assert cname in hs and csname in cs, (cname, csname)
        hf, cf = hs[cname], cs[csname]
        assert len(hf) == len(cf) == len(ctype._fields_), (cname, hf, cf)
        for (hn, ht, hc), (cn, ct, cc), (pn, pt) in zip(hf, cf, ctype._fields_):
            assert hn.replace("_", "").lower() == cn.lower(), (cname, hn, cn)       # same field, same position
            assert pn == hn, (cname, pn, hn)                                            # ctypes view uses the header's names
            if ht == "rtc_hit":
                assert ct == "RtcHit" and pt is N.Hit
                continue
            hsize = C_SIZES[ht] * hc
            csize = (8 if ct == "ptr" else CS_SIZES[ct]) * cc
            assert hsize == csize == C.sizeof(pt), (cname, hn, hsize, csize, C.sizeof(pt))
    # natural alignment makes LayoutKind.Sequential agree with the C layout when sizes and order agree; spot-check totals
    assert (C.sizeof(N.Hit), C.sizeof(N.DebugRay), C.sizeof(N.BvhNode), C.sizeof(N.Camera), C.sizeof(N.Params), C.sizeof(N.Stats)) == (64, 80, 64, 176, 56, 128)


def test_shim_has_the_reference_surface():
    """Members of FullRaytracer that MainWindow.cs / RayInspector.cs use (FullRaytracer.cs:35-36,61-66,131,179,243,375-416;
    DebugRaycaster.cs:118-165,217) must exist on the shim with the reference's signatures."""
    for pat in (r"public GpuFullRaytracer\(Scene scene, int threads, Action<GpuFullRaytracer, string, double, Bitmap> updateStatus,\s*Action<GpuFullRaytracer, Bitmap> updateDebug\)",
                r"public Scene Scene;", r"public double Exposure", r"public readonly GpuDebugPathtracer DebugPathtracer;",
                r"public readonly GpuDebugRaycaster DebugRaycaster;", r"public void Start\(\)", r"public void Stop\(\)",
                r"public void Pause\(\)", r"public void Resume\(\)", r"public void QueueUpdate\(\)", r"public void QueueDebugUpdate\(\)",
                r"public bool IsRunning", r"public bool IsPaused", r"public bool IsStopping", r"public SampleSet GetSampleSet\(int x, int y\)",
                r"public Bitmap GetBitmap\(\)", r"public Raytracer.DebugRay\[\] GetDebugTrace\(int x, int y\)",
                r"class GpuDebugRaycaster : DebugRaycaster", r"public new void SetMode\(DisplayMode mode\)",
                r"public new bool SetDisplayOnly\(object item\)", r"public new void ClearDisplayOnly\(\)", r"public new Bitmap RenderDebug\(\)"):
        assert re.search(pat, SHIM), pat
    used = set(re.findall(r"RtcoreNative\.(rtc_\w+)", SHIM))
    assert used <= set(cs_imports()), used - set(cs_imports())
    assert {"rtc_debug_trace", "rtc_debug_raycast", "rtc_read_pixel", "rtc_tonemap_argb", "rtc_render", "rtc_get_stats"} <= used
