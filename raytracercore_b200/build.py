"""In-tree build of librtcore_b200.so (CUDA kernels + C ABI + C++ host layer) for sm_100a.

`python -m raytracercore_b200.build` or `__graft_entry__.build()`. nvcc cross-compiles without a GPU. The two
kernel translation units differ only in -fmad: the f64 parity kernels are built with -fmad=false so that the
only fused operations are the explicit fma() calls mirroring the reference's Fma.* intrinsics.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "librtcore_b200.so")
# the scene half of the host mirror alone (SceneLoader, synthetic scenes, BVH builder; no CUDA): what bench.py's
# `--impl reference` arm loads, so that the CPU process maps no CUDA library
HOST_LIB = os.path.join(HERE, "librtcore_host.so")
HOST_ONLY_SOURCES = ("scene.cpp", "scene_loader.cpp", "bvh_builder.cpp", "host_api.cpp")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_COMMON = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build step failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build_variant(name, defines):
    """Development aid: build lib variant `name` with extra -D flags (kernel tuning A/B runs via RTC_B200_LIB)."""
    global OBJ, LIB
    old = (OBJ, LIB)
    OBJ = os.path.join(HERE, "_obj_" + name)
    LIB = os.path.join(HERE, "librtcore_b200_%s.so" % name)
    extra = os.environ.get("RTC_EXTRA_NVCC", "")
    os.environ["RTC_EXTRA_NVCC"] = " ".join(defines)
    try:
        return build(force=True)
    finally:
        os.environ["RTC_EXTRA_NVCC"] = extra
        OBJ, LIB = old


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers += [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".h")]
    headers += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    all_src = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu")]
    all_src += [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".cpp")]
    if not force and not _newer(LIB, all_src + headers) and os.path.exists(HOST_LIB):
        return LIB  # up to date (also the case on the GPU box, where the prebuilt library travels with the snapshot)
    objs = []
    units = [
        ("kernels_f32.cu", NVCC_COMMON),
        ("kernels_f64.cu", NVCC_COMMON + ["-fmad=false"]),
        ("rtc_api.cu", NVCC_COMMON),
        ("bvh_build.cu", NVCC_COMMON),
        ("reorder.cu", NVCC_COMMON),
        # Scene.Prepare on the device: bit-identical to the host builder / flatten, hence no FMA contraction here either
        ("prepare_device.cu", NVCC_COMMON + ["-fmad=false"]),
    ]
    for src, flags in units:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src + ".o")
        if force or _newer(o, [s] + headers):
            extra = os.environ.get("RTC_EXTRA_NVCC", "").split()
            out = _run([nvcc] + flags + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])
            if verbose:
                print(out)
        objs.append(o)
    cxx = os.environ.get("CXX", "g++")
    for src in sorted(f for f in os.listdir(HOST) if f.endswith(".cpp")):
        s = os.path.join(HOST, src)
        o = os.path.join(OBJ, src + ".o")
        if force or _newer(o, [s] + headers):
            _run([cxx, "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-Wall", "-pthread", "-c", s, "-o", o])
        objs.append(o)
    if force or _newer(LIB, objs):
        _run([nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-ldl", "-lpthread"])
    host_objs = [os.path.join(OBJ, f + ".o") for f in HOST_ONLY_SOURCES]
    if LIB.endswith("librtcore_b200.so") and (force or _newer(HOST_LIB, host_objs)):
        _run([cxx, "-shared", "-o", HOST_LIB] + host_objs + ["-lpthread"])
    return LIB


def build_oracle(force=False):
    """Builds oracle/librtc_oracle.so (test infrastructure; never loaded by the product)."""
    odir = os.path.join(ROOT, "oracle")
    if force:
        subprocess.run(["make", "-C", odir, "clean"], stdout=subprocess.DEVNULL)
    _run(["make", "-C", odir])
    return os.path.join(odir, "librtc_oracle.so")


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_oracle())
