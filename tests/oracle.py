"""ctypes view of oracle/librtc_oracle.so — the CPU restatement of the reference render path. Test infrastructure:
imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os

import numpy as np

from raytracercore_b200 import _native as N
from raytracercore_b200.renderer import HIT_DT, RAY_DT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "librtc_oracle.so")


def _load():
    if not os.path.exists(LIB_PATH):
        from raytracercore_b200 import build as B
        B.build_oracle()
    lib = C.CDLL(LIB_PATH)
    P = C.c_void_p
    lib.orc_scene_create.restype = P
    lib.orc_scene_create.argtypes = [C.POINTER(N.SceneDesc), C.c_int32, C.POINTER(N.BvhNode), C.c_int32, C.POINTER(N.Camera), C.POINTER(N.Params)]
    lib.orc_scene_destroy.argtypes = [P]
    lib.orc_set_camera.argtypes = [P, C.POINTER(N.Camera)]
    lib.orc_set_params.argtypes = [P, C.POINTER(N.Params)]
    lib.orc_trace_closest.restype = C.c_int64
    lib.orc_trace_closest.argtypes = [P, C.c_int64, P, P, P, C.c_int, C.c_int, C.c_int]
    lib.orc_camera_rays.argtypes = [P, C.c_int64, P, P, P]
    lib.orc_render.restype = C.c_uint64
    lib.orc_render.argtypes = [P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, C.c_uint32, C.c_int, P, P, P]
    lib.orc_render_lattice.restype = C.c_uint64
    lib.orc_render_lattice.argtypes = [P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, C.c_uint32, C.c_int, P, P, P]
    lib.orc_render_samples.argtypes = [P, C.c_uint32, C.c_int, P]
    lib.orc_debug_trace.argtypes = [P, C.c_int32, C.c_int32, C.c_uint32, C.c_int32, C.POINTER(N.DebugRay), C.POINTER(C.c_int32)]
    lib.orc_debug_raycast.argtypes = [P, C.c_int32, P]
    lib.orc_set_selfhit_mode.argtypes = [C.c_int]
    lib.orc_dump_path_rays.restype = C.c_int64
    lib.orc_dump_path_rays.argtypes = [P, C.c_int64, P, P, C.c_int, C.c_int64, P, P, P, P]
    lib.orc_tonemap.argtypes = [C.c_int32, C.c_int32, P, P, P, C.c_double, C.POINTER(C.c_double), C.c_double, P]
    lib.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.orc_uniforms.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]
    lib.orc_create_horizon.argtypes = [C.POINTER(C.c_double), C.c_double, C.c_double, C.POINTER(C.c_double)]
    lib.orc_aabb_intersect.restype = C.c_int
    lib.orc_aabb_intersect.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(N.Ray), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    return lib


lib = _load()
NTHREADS = os.cpu_count() or 1


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleScene:
    """The reference's Scene + Raytracer over the same flattened primitives and the same BVH the GPU is given."""

    def __init__(self, scene, seed=1, camera=None, bvh=None):
        self.scene = scene
        d = scene.desc()
        nodes, n, root = bvh if bvh is not None else scene.bvh()
        self.params = scene.params(seed)
        self.cam = scene.camera(camera)
        self._h = C.c_void_p(lib.orc_scene_create(C.byref(d), n, nodes, root, C.byref(self.cam), C.byref(self.params)))
        self.width, self.height = self.params.width, self.params.height

    def close(self):
        if self._h:
            lib.orc_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, p):
        self.params = p
        self.width, self.height = p.width, p.height
        lib.orc_set_params(self._h, C.byref(p))

    def set_camera(self, c):
        self.cam = c
        lib.orc_set_camera(self._h, C.byref(c))

    def trace_closest(self, rays, skip=None, mode=0, check_both=False, threads=NTHREADS):
        rays = np.ascontiguousarray(rays, dtype=RAY_DT)
        out = np.zeros(len(rays), dtype=HIT_DT)
        sp = None
        if skip is not None:
            skip = np.ascontiguousarray(skip, dtype=HIT_DT)
            sp = _ptr(skip)
        diff = lib.orc_trace_closest(self._h, len(rays), _ptr(rays), sp, _ptr(out), mode, 1 if check_both else 0, threads)
        return (out, diff) if check_both else out

    def camera_rays(self, xy, sample):
        xy = np.ascontiguousarray(xy, dtype=np.int32).reshape(-1, 2)
        sample = np.ascontiguousarray(sample, dtype=np.uint32)
        out = np.zeros(len(xy), dtype=RAY_DT)
        lib.orc_camera_rays(self._h, len(xy), _ptr(xy), _ptr(sample), _ptr(out))
        return out

    def render(self, first_sample, n_samples, rect=None, threads=NTHREADS, accum=None):
        x0, y0, x1, y1 = rect if rect is not None else (0, 0, self.width, self.height)
        if accum is None:
            accum = (np.zeros((self.height, self.width, 3)), np.zeros((self.height, self.width), np.uint32),
                     np.zeros((self.height, self.width), np.uint32))
        rgb, s, m = accum
        rays = lib.orc_render(self._h, x0, y0, x1, y1, first_sample, n_samples, threads, _ptr(rgb), _ptr(s), _ptr(m))
        return rgb, s, m, int(rays)

    def render_lattice(self, stride, offset, first_sample, n_samples, threads=NTHREADS, accum=None):
        """orc_render over the lattice of pixels (offset[0] + i stride[0], offset[1] + j stride[1]) of the whole frame."""
        if accum is None:
            accum = (np.zeros((self.height, self.width, 3)), np.zeros((self.height, self.width), np.uint32),
                     np.zeros((self.height, self.width), np.uint32))
        rgb, s, m = accum
        rays = lib.orc_render_lattice(self._h, stride[0], stride[1], offset[0], offset[1], first_sample, n_samples, threads,
                                      _ptr(rgb), _ptr(s), _ptr(m))
        return rgb, s, m, int(rays)

    def render_samples(self, sample, threads=NTHREADS):
        out = np.zeros((self.height, self.width, 3))
        lib.orc_render_samples(self._h, sample, threads, _ptr(out))
        return out

    def dump_path_rays(self, xy, sample, threads=NTHREADS):
        """Every Scene.RayTrace call GetColor makes for the paths (x, y, sample): (rays, skip hits, oracle hits, bounce)."""
        xy = np.ascontiguousarray(xy, dtype=np.int32).reshape(-1, 2)
        sample = np.ascontiguousarray(sample, dtype=np.uint32)
        cap = len(xy) * (self.params.recursion + 1)
        rays, skip, hits = np.zeros(cap, RAY_DT), np.zeros(cap, HIT_DT), np.zeros(cap, HIT_DT)
        bounce = np.zeros(cap, np.int32)
        n = lib.orc_dump_path_rays(self._h, len(xy), _ptr(xy), _ptr(sample), threads, cap, _ptr(rays), _ptr(skip), _ptr(hits), _ptr(bounce))
        assert 0 <= n <= cap
        return rays[:n], skip[:n], hits[:n], bounce[:n]

    def debug_raycast(self, mode):
        out = np.zeros((self.height, self.width), dtype=np.int32)
        lib.orc_debug_raycast(self._h, mode, _ptr(out))
        return out

    def debug_trace(self, x, y, sample, capacity=64):
        buf = (N.DebugRay * capacity)()
        n = C.c_int32()
        lib.orc_debug_trace(self._h, x, y, sample, capacity, buf, C.byref(n))
        return [buf[i] for i in range(n.value)]


def set_selfhit_mode(mode):
    """0: the reference's self-hit rule; 1: the f32 mode's documented rule restated in f64 (see rtc_oracle.h)."""
    lib.orc_set_selfhit_mode(int(mode))


def tonemap(rgb, samples, misses, exposure=1.0, back=(0.0, 0.0, 0.0), back_a=0.0):
    h, w = samples.shape
    out = np.zeros((h, w), dtype=np.uint32)
    rgb = np.ascontiguousarray(rgb, dtype=np.float64)
    samples = np.ascontiguousarray(samples, dtype=np.uint32)
    misses = np.ascontiguousarray(misses, dtype=np.uint32)
    b = (C.c_double * 3)(*back)
    lib.orc_tonemap(w, h, _ptr(rgb), _ptr(samples), _ptr(misses), exposure, b, back_a, _ptr(out))
    return out


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib.orc_philox4x32_10(c, k, o)
    return tuple(o)


def uniforms(seed, pixel, sample, stage, block):
    o = (C.c_double * 2)()
    lib.orc_uniforms(seed, pixel, sample, stage, block, o)
    return o[0], o[1]


def create_horizon(pole, z, theta):
    p = (C.c_double * 3)(*pole)
    o = (C.c_double * 3)()
    lib.orc_create_horizon(p, z, theta, o)
    return np.array(o[:])


def aabb_intersect(bmin, bmax, origin, direction):
    r = N.Ray()
    r.origin[:] = origin
    r.dir[:] = direction
    n = C.c_double()
    f = C.c_double()
    ok = lib.orc_aabb_intersect((C.c_double * 3)(*bmin), (C.c_double * 3)(*bmax), C.byref(r), C.byref(n), C.byref(f))
    return bool(ok), n.value, f.value
