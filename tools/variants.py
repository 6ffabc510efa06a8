"""Development aid: build kernel variants of librtcore_b200 that differ only in -D flags of kernels_f32.cu (the other
objects are shared with the main build), for A/B runs on the GPU box via RTC_B200_LIB.

  python tools/variants.py name1:"-DRTC_TRACE_MIN_BLOCKS=5" name2:"-DRTC_TRACE_THREADS=64 -DRTC_TRACE_MIN_BLOCKS=11"
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytracercore_b200 import build as B  # noqa: E402


def build_variant(name, flags):
    B.build()  # main objects up to date
    vdir = os.path.join(B.HERE, "_variants")
    os.makedirs(vdir, exist_ok=True)
    obj = os.path.join(vdir, "kernels_f32_%s.o" % name)
    lib = os.path.join(vdir, "librtcore_b200_%s.so" % name)
    nvcc = B._nvcc()
    out = B._run([nvcc] + B.NVCC_COMMON + flags.split() + ["-Xptxas", "-v", "-c", os.path.join(B.CSRC, "kernels_f32.cu"), "-o", obj])
    lines = out.splitlines()
    for i, l in enumerate(lines):
        if "k_trace_q8ILb0" in l and "Compiling" in l:
            print("  [%s] %s | %s" % (name, lines[i + 2].strip(), lines[i + 3].strip()))
    objs = [os.path.join(B.OBJ, f) for f in sorted(os.listdir(B.OBJ)) if f.endswith(".o") and f != "kernels_f32.cu.o"] + [obj]
    B._run([nvcc] + B.ARCH + ["-shared", "-o", lib] + objs + ["-ldl", "-lpthread"])
    return lib


if __name__ == "__main__":
    for spec in sys.argv[1:]:
        name, _, flags = spec.partition(":")
        print(build_variant(name, flags))
