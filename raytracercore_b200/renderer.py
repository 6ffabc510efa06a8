"""Context (the kernel ABI, include/rtcore_b200.h) and FullRaytracer (mirror of the reference's
Raytracing/FullRaytracer.cs) — thin Python views used by the tests, bench.py and the multi-GPU driver."""
import ctypes as C

import numpy as np

from . import _native as N
from .scene import Scene

RAY_DT = np.dtype([("origin", "<f8", 3), ("dir", "<f8", 3)])
HIT_DT = np.dtype([("prim", "<i4"), ("inside", "<i4"), ("t", "<f8"), ("position", "<f8", 3), ("normal", "<f8", 3)])
assert RAY_DT.itemsize == C.sizeof(N.Ray) and HIT_DT.itemsize == C.sizeof(N.Hit)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Baked:
    def __init__(self, h):
        self._h = h

    @property
    def nbytes(self):
        return int(N.lib.rtc_baked_bytes(self._h))

    def segments(self):
        """{name: bytes} of the image's arrays (copies)."""
        out = {}
        for i, name in enumerate(N.BAKED_SEGMENT_NAMES):
            p, nb = C.c_void_p(), C.c_int64()
            rc = N.lib.rtc_baked_segment(self._h, i, C.byref(p), C.byref(nb))
            if rc != N.RTC_OK:
                raise N.RtcError(rc, "rtc_baked_segment")
            out[name] = C.string_at(p, nb.value) if nb.value else b""
        return out

    def close(self):
        if self._h:
            N.lib.rtc_baked_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One rtc_ctx: one GPU, one caller at a time."""

    def __init__(self, device=0, precision=N.RTC_F32):
        h = C.c_void_p()
        rc = N.lib.rtc_create(device, precision, C.byref(h))
        if rc != N.RTC_OK:
            msg = N.lib.rtc_last_error(None)
            raise N.RtcError(rc, msg.decode() if msg else "")
        self._h = h
        self.precision = precision
        self.width = self.height = 0
        self._keep = None

    def close(self):
        if self._h:
            N.lib.rtc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        N.check(self._h, rc)

    def set_option(self, opt, value):
        self._ck(N.lib.rtc_set_option(self._h, opt, int(value)))

    def set_stream(self, cuda_stream):
        """cuda_stream: integer cudaStream_t handle (e.g. torch.cuda.current_stream().cuda_stream) or None."""
        self._ck(N.lib.rtc_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None))

    # -- scene hand-over ---------------------------------------------------------------------------------
    def upload_scene(self, scene_or_desc):
        d = scene_or_desc.desc() if isinstance(scene_or_desc, Scene) else scene_or_desc
        self._ck(N.lib.rtc_upload_scene(self._h, C.byref(d)))

    def upload_bvh(self, nodes, n_nodes, root):
        self._ck(N.lib.rtc_upload_bvh(self._h, n_nodes, nodes, root))

    def build_bvh(self, device=False, radius=0):
        """Host binned-SAH build, or (device=True) the clustering build on the GPU; returns its number of rounds."""
        if not device:
            self._ck(N.lib.rtc_build_bvh(self._h))
            return 0
        rounds = C.c_int32()
        self._ck(N.lib.rtc_build_bvh_device(self._h, radius, C.byref(rounds)))
        return rounds.value

    def create_horizon(self, pole_z_theta):
        """Vec4D.CreateHorizon on the device for an (n, 5) array of (pole.xyz, z, theta); returns (n, 3)."""
        a = np.ascontiguousarray(pole_z_theta, dtype=np.float64).reshape(-1, 5)
        out = np.zeros((len(a), 3), np.float64)
        self._ck(N.lib.rtc_debug_create_horizon(self._h, len(a), _ptr(a), _ptr(out)))
        return out

    def prepare_device(self, builder=N.RTC_BUILDER_SAH, radius=0):
        """Scene.Prepare wholly on the GPU (tree + device layout); returns the rtc_prepare_stats."""
        st = N.PrepareStats()
        self._ck(N.lib.rtc_prepare_device(self._h, builder, radius, C.byref(st)))
        return st

    def get_bvh(self):
        n = C.c_int32()
        root = C.c_int32()
        self._ck(N.lib.rtc_get_bvh_size(self._h, C.byref(n), C.byref(root)))
        nodes = (N.BvhNode * n.value)()
        self._ck(N.lib.rtc_get_bvh(self._h, n.value, nodes))
        return nodes, n.value, root.value

    def bake(self):
        """Host-resident (pinned) image of the current device scene; re-upload with upload_baked()."""
        h = C.c_void_p()
        self._ck(N.lib.rtc_bake(self._h, C.byref(h)))
        return Baked(h)

    def upload_baked(self, baked):
        self._ck(N.lib.rtc_upload_baked(self._h, baked._h))

    def set_camera(self, cam):
        self._ck(N.lib.rtc_set_camera(self._h, C.byref(cam)))

    def set_params(self, par):
        self._ck(N.lib.rtc_set_params(self._h, C.byref(par)))
        self.width, self.height = par.width, par.height

    def load(self, scene, seed=1, camera=None, use_scene_bvh=True, device_bvh=False, radius=0, device_prepare=None):
        """Scene.Prepare + FullRaytracer.Start's set-up (FullRaytracer.cs:253-269) in one call. device_bvh: build the
        tree on the GPU (rtc_build_bvh_device) instead of taking the host scene's accelerator. device_prepare: a builder
        (RTC_BUILDER_SAH / RTC_BUILDER_PLOC): tree and device layout both made on the GPU (rtc_prepare_device)."""
        self.upload_scene(scene)
        self.prepare_stats = None
        if device_prepare is not None:
            self.prepare_stats = self.prepare_device(device_prepare, radius)
        elif device_bvh:
            self.build_bvh(device=True, radius=radius)
        elif use_scene_bvh:
            nodes, n, root = scene.bvh()
            self.upload_bvh(nodes, n, root)
        else:
            self.build_bvh()
        self.set_params(scene.params(seed))
        self.set_camera(scene.camera(camera))

    # -- hot path ----------------------------------------------------------------------------------------
    def trace_closest(self, rays, skip=None):
        rays = np.ascontiguousarray(rays, dtype=RAY_DT)
        out = np.zeros(len(rays), dtype=HIT_DT)
        sp = None
        if skip is not None:
            skip = np.ascontiguousarray(skip, dtype=HIT_DT)
            assert len(skip) == len(rays)
            sp = _ptr(skip)
        self._ck(N.lib.rtc_trace_closest(self._h, len(rays), _ptr(rays), sp, _ptr(out)))
        return out

    def camera_rays(self, xy, sample):
        xy = np.ascontiguousarray(xy, dtype=np.int32).reshape(-1, 2)
        sample = np.ascontiguousarray(sample, dtype=np.uint32)
        out = np.zeros(len(xy), dtype=RAY_DT)
        self._ck(N.lib.rtc_camera_rays(self._h, len(xy), _ptr(xy), _ptr(sample), _ptr(out)))
        return out

    def render(self, first_sample, n_samples, rect=None):
        x0, y0, x1, y1 = rect if rect is not None else (0, 0, self.width, self.height)
        self._ck(N.lib.rtc_render(self._h, x0, y0, x1, y1, first_sample, n_samples))

    def render_read(self, first_sample, n_samples, out=None):
        """rtc_render over the whole image + rtc_read_accum, band read-back overlapped with rendering.
        out: optional (rgb f64 [h,w,3], samples u32 [h,w], misses u32 [h,w]) arrays or raw pointers (e.g. pinned)."""
        if out is None:
            rgb = np.empty((self.height, self.width, 3), np.float64)
            s = np.empty((self.height, self.width), np.uint32)
            m = np.empty((self.height, self.width), np.uint32)
            self._ck(N.lib.rtc_render_read(self._h, first_sample, n_samples, _ptr(rgb), _ptr(s), _ptr(m)))
            return rgb, s, m
        ptrs = [(_ptr(a) if isinstance(a, np.ndarray) else a) for a in out]
        self._ck(N.lib.rtc_render_read(self._h, first_sample, n_samples, *ptrs))
        return out

    def sync(self):
        self._ck(N.lib.rtc_sync(self._h))

    def clear_accum(self):
        self._ck(N.lib.rtc_clear_accum(self._h))

    def read_accum(self, out=None):
        """out: optional (rgb f64 [h,w,3], samples u32 [h,w], misses u32 [h,w]) arrays or raw pointers (e.g. pinned)."""
        assert self.width * self.height > 0
        if out is None:
            rgb = np.zeros((self.height, self.width, 3), dtype=np.float64)
            s = np.zeros((self.height, self.width), dtype=np.uint32)
            m = np.zeros((self.height, self.width), dtype=np.uint32)
            self._ck(N.lib.rtc_read_accum(self._h, _ptr(rgb), _ptr(s), _ptr(m)))
            return rgb, s, m
        ptrs = [C.c_void_p(o) if isinstance(o, int) else _ptr(o) for o in out]
        self._ck(N.lib.rtc_read_accum(self._h, *ptrs))
        return out

    def write_accum(self, rgb, samples, misses):
        rgb = np.ascontiguousarray(rgb, dtype=np.float64)
        samples = np.ascontiguousarray(samples, dtype=np.uint32)
        misses = np.ascontiguousarray(misses, dtype=np.uint32)
        self._ck(N.lib.rtc_write_accum(self._h, _ptr(rgb), _ptr(samples), _ptr(misses)))

    def accum_device_ptrs(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._ck(N.lib.rtc_accum_device_ptrs(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def tonemap(self, exposure=1.0, back=(0.0, 0.0, 0.0), back_a=0.0):
        out = np.zeros((self.height, self.width), dtype=np.uint32)
        b = (C.c_double * 3)(*back)
        self._ck(N.lib.rtc_tonemap_argb(self._h, exposure, b, back_a, _ptr(out)))
        return out

    def read_pixel(self, x, y):
        """FullRaytracer.GetSampleSet(x, y): (rgb sum, samples, misses) of one pixel, on the read-out stream."""
        rgb = (C.c_double * 3)()
        s, m = C.c_uint32(), C.c_uint32()
        self._ck(N.lib.rtc_read_pixel(self._h, x, y, rgb, C.byref(s), C.byref(m)))
        return tuple(rgb[:]), s.value, m.value

    def render_samples(self, sample):
        out = np.zeros((self.height, self.width, 3), dtype=np.float64)
        self._ck(N.lib.rtc_render_samples(self._h, sample, _ptr(out)))
        return out

    def debug_trace(self, x, y, sample, capacity=64):
        buf = (N.DebugRay * capacity)()
        n = C.c_int32()
        self._ck(N.lib.rtc_debug_trace(self._h, x, y, sample, capacity, buf, C.byref(n)))
        return [buf[i] for i in range(n.value)]

    def debug_raycast(self, mode):
        """DebugRaycaster overlay query: mode 0 = primitive ids (-1 none), 1 = BVH box-intersection counts."""
        out = np.zeros((self.height, self.width), dtype=np.int32)
        self._ck(N.lib.rtc_debug_raycast(self._h, mode, _ptr(out)))
        return out

    def debug_raycast_selection(self, prim_ids):
        """DebugRaycaster's Selection mode over primitives: per pixel the ID of the nearest selected primitive, or -1."""
        ids = np.ascontiguousarray(prim_ids, dtype=np.int32)
        out = np.zeros((self.height, self.width), dtype=np.int32)
        self._ck(N.lib.rtc_debug_raycast_selection(self._h, len(ids), _ptr(ids), _ptr(out)))
        return out

    def stats(self):
        s = N.Stats()
        self._ck(N.lib.rtc_get_stats(self._h, C.byref(s)))
        return s

    def reset_stats(self):
        self._ck(N.lib.rtc_reset_stats(self._h))

    # -- multi-GPU ---------------------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(128)
        rc = N.lib.rtc_comm_unique_id(buf)
        if rc:
            raise N.RtcError(rc, (N.lib.rtc_last_error(None) or b"").decode())
        return buf.raw

    def comm_init(self, nranks, rank, uid):
        buf = C.create_string_buffer(uid, 128)
        self._ck(N.lib.rtc_comm_init(self._h, nranks, rank, buf))

    def reduce_accum(self, root=0):
        self._ck(N.lib.rtc_reduce_accum(self._h, root))

    def bcast_scene(self, root=0):
        """Collective: every rank receives root's device scene over NVLink (no rtc_upload_scene / rtc_upload_bvh on the others)."""
        self._ck(N.lib.rtc_bcast_scene(self._h, root))


class FullRaytracer:
    """Mirror of FullRaytracer (Raytracing/FullRaytracer.cs): Start() blocks; Stop/Pause/Resume from other threads."""

    def __init__(self, scene, device=0, precision=N.RTC_F32, seed=1, update_status=None):
        self.Scene = scene
        self._cb_user = update_status

        def _cb(_user, text, progress):
            if self._cb_user:
                self._cb_user(self, text.decode(), progress)

        self._cb = N.STATUS_FN(_cb)
        err = C.create_string_buffer(512)
        h = N.lib.rtcs_raytracer_create(scene._h, device, precision, int(seed), self._cb, None, err, len(err))
        if not h:
            raise N.RtcError(N.RTC_ERR_CUDA, err.value.decode())
        self._h = C.c_void_p(h)
        self._exposure = 1.0

    def close(self):
        if self._h:
            N.lib.rtcs_raytracer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def Start(self, samples_per_pass=1, max_samples=0):
        rc = N.lib.rtcs_raytracer_start(self._h, samples_per_pass, max_samples)
        if rc:
            raise N.RtcError(rc, N.lib.rtcs_raytracer_last_error(self._h).decode())

    def Stop(self):
        N.lib.rtcs_raytracer_stop(self._h)

    def Pause(self):
        N.lib.rtcs_raytracer_pause(self._h)

    def Resume(self):
        N.lib.rtcs_raytracer_resume(self._h)

    @property
    def IsRunning(self):
        return bool(N.lib.rtcs_raytracer_is_running(self._h))

    @property
    def IsPaused(self):
        return bool(N.lib.rtcs_raytracer_is_paused(self._h))

    @property
    def IsStopping(self):
        return bool(N.lib.rtcs_raytracer_is_stopping(self._h))

    @property
    def Exposure(self):
        return self._exposure

    @Exposure.setter
    def Exposure(self, v):
        self._exposure = float(v)
        N.lib.rtcs_raytracer_set_exposure(self._h, self._exposure)

    def GetSampleSet(self, x, y):
        rgb = (C.c_double * 3)()
        s = C.c_uint32()
        m = C.c_uint32()
        rc = N.lib.rtcs_raytracer_get_sample_set(self._h, x, y, rgb, C.byref(s), C.byref(m))
        if rc:
            raise N.RtcError(rc, N.lib.rtcs_raytracer_last_error(self._h).decode())
        return tuple(rgb[:]), s.value, m.value

    def GetBitmap(self):
        g = self.Scene.globals()
        out = np.zeros((g.height, g.width), dtype=np.uint32)
        rc = N.lib.rtcs_raytracer_get_bitmap(self._h, _ptr(out))
        if rc == N.RTC_ERR_STATE:
            return None  # GetBitmap returns null before the first Start (FullRaytracer.cs:184-185)
        if rc:
            raise N.RtcError(rc, N.lib.rtcs_raytracer_last_error(self._h).decode())
        return out
