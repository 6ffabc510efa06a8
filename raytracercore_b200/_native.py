"""ctypes binding of librtcore_b200.so (include/rtcore_b200.h + include/rtcore_host.h).

The library is the product: there is no Python or CPU fallback. If the shared object is missing, importing this
module raises (run `python -m raytracercore_b200.build` or `__graft_entry__.build()`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RTC_B200_HOST_ONLY=1 (bench.py --impl reference): only the scene half of the host mirror (librtcore_host.so: SceneLoader,
# synthetic scenes, BVH builder) is loaded -- the process maps no CUDA library and the rtc_* / rtcs_raytracer_* calls do not exist
HOST_ONLY = os.environ.get("RTC_B200_HOST_ONLY") == "1"
LIB_PATH = os.path.join(_HERE, "librtcore_host.so") if HOST_ONLY else (
    os.environ.get("RTC_B200_LIB") or os.path.join(_HERE, "librtcore_b200.so"))  # env override: kernel-variant A/B runs

RTC_OK, RTC_ERR_INVALID, RTC_ERR_CUDA, RTC_ERR_STATE, RTC_ERR_NOMEM, RTC_ERR_UNSUPPORTED, RTC_ERR_NCCL = range(7)
RTC_F32, RTC_F64 = 0, 1
RTC_KIND_TRIANGLE, RTC_KIND_SPHERE, RTC_KIND_PLANE = 0, 1, 2
RTC_FLAG_MIRROR, RTC_FLAG_TWOSIDED, RTC_FLAG_INVERT, RTC_FLAG_TRANSFORMED, RTC_FLAG_VNORMALS = 1, 2, 4, 8, 16
RTC_CAMERA_FRUSTUM, RTC_CAMERA_ORTHO = 0, 1
RTC_GEOM_STRIDE, RTC_MATERIAL_STRIDE, RTC_XFORM_STRIDE = 12, 14, 48
RTC_K_RAYGEN, RTC_K_TRACE, RTC_K_SHADE, RTC_K_COMPACT, RTC_K_ACCUMULATE, RTC_K_COUNT = 0, 1, 2, 3, 4, 5
RTC_OPT_KERNEL_TIMING, RTC_OPT_COUNTERS, RTC_OPT_MAX_PATHS, RTC_OPT_WAVES, RTC_OPT_REORDER = 1, 2, 3, 4, 5
RTC_BUILDER_SAH, RTC_BUILDER_PLOC = 0, 1
RTC_BAKED_SEGMENTS = 10
BAKED_SEGMENT_NAMES = ("nodes", "qnodes", "unbounded", "prims", "mats", "xforms", "aux", "prim_id", "id_to_slot", "sgeom")
KERNEL_NAMES = ("raygen", "trace", "shade", "compact", "accumulate")
BOUNCE_TYPES = ("Skipped", "Diffuse", "Specular", "SpecularFail", "Transmitted", "Emission", "PureBlack",
                "RecursionComplete", "Missed", "Debug")


class SceneDesc(C.Structure):
    _fields_ = [("n_prims", C.c_int32), ("n_xforms", C.c_int32), ("kind", C.POINTER(C.c_uint8)),
                ("flags", C.POINTER(C.c_uint8)), ("geom", C.POINTER(C.c_double)), ("xform", C.POINTER(C.c_int32)),
                ("xforms", C.POINTER(C.c_double)), ("material", C.POINTER(C.c_double))]


class BvhNode(C.Structure):
    _fields_ = [("bmin", C.c_double * 3), ("bmax", C.c_double * 3), ("left", C.c_int32), ("right", C.c_int32),
                ("prim", C.c_int32), ("pad", C.c_int32)]


class Camera(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pad", C.c_int32), ("position", C.c_double * 3), ("look", C.c_double * 3),
                ("side", C.c_double * 3), ("up", C.c_double * 3), ("w2", C.c_double), ("h2", C.c_double),
                ("tan_fov_x2", C.c_double), ("tan_fov_y2", C.c_double), ("h_mult", C.c_double), ("v_mult", C.c_double),
                ("image_plane", C.c_double), ("dof_amount", C.c_double), ("focal_length", C.c_double)]


class Params(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("recursion", C.c_int32), ("debug_geom", C.c_int32),
                ("ambient", C.c_double * 3), ("air_ior", C.c_double), ("seed", C.c_uint64)]


class PrepareStats(C.Structure):
    _fields_ = [("boxes_ms", C.c_double), ("build_ms", C.c_double), ("flatten_ms", C.c_double), ("total_ms", C.c_double),
                ("build_levels", C.c_int32), ("wide_depth", C.c_int32), ("n_wide_nodes", C.c_int32), ("n_bounded", C.c_int32)]


class Ray(C.Structure):
    _fields_ = [("origin", C.c_double * 3), ("dir", C.c_double * 3)]


class Hit(C.Structure):
    _fields_ = [("prim", C.c_int32), ("inside", C.c_int32), ("t", C.c_double), ("position", C.c_double * 3),
                ("normal", C.c_double * 3)]


class DebugRay(C.Structure):
    _fields_ = [("hit", Hit), ("type", C.c_int32), ("pad", C.c_int32), ("fresnel_ratio", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays", C.c_uint64), ("launches", C.c_uint64 * RTC_K_COUNT),
                ("ms", C.c_double * RTC_K_COUNT), ("nodes_visited", C.c_uint64), ("prims_tested", C.c_uint64),
                ("node_steps", C.c_uint64), ("leaf_steps", C.c_uint64)]


class Globals(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("recursion", C.c_int32), ("debug_geom", C.c_int32),
                ("n_cameras", C.c_int32), ("current_camera", C.c_int32), ("n_prims", C.c_int32), ("pad", C.c_int32),
                ("background", C.c_double * 3), ("background_alpha", C.c_double), ("ambient", C.c_double * 3),
                ("air_ior", C.c_double)]


STATUS_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p, C.c_double)

# name -> (restype, argtypes); every symbol include/*.h declares
_P = C.c_void_p
SIGNATURES = {
    # rtcore_b200.h
    "rtc_abi_version": (C.c_int, []),
    "rtc_device_count": (C.c_int, []),
    "rtc_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(_P)]),
    "rtc_destroy": (None, [_P]),
    "rtc_last_error": (C.c_char_p, [_P]),
    "rtc_set_option": (C.c_int, [_P, C.c_int, C.c_int64]),
    "rtc_set_stream": (C.c_int, [_P, _P]),
    "rtc_upload_scene": (C.c_int, [_P, C.POINTER(SceneDesc)]),
    "rtc_upload_bvh": (C.c_int, [_P, C.c_int32, C.POINTER(BvhNode), C.c_int32]),
    "rtc_build_bvh": (C.c_int, [_P]),
    "rtc_bake": (C.c_int, [_P, C.POINTER(_P)]),
    "rtc_upload_baked": (C.c_int, [_P, _P]),
    "rtc_baked_bytes": (C.c_int64, [_P]),
    "rtc_baked_free": (None, [_P]),
    "rtc_baked_segment": (C.c_int, [_P, C.c_int32, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "rtc_prepare_device": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(PrepareStats)]),
    "rtc_get_bvh_size": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "rtc_get_bvh": (C.c_int, [_P, C.c_int32, C.POINTER(BvhNode)]),
    "rtc_set_camera": (C.c_int, [_P, C.POINTER(Camera)]),
    "rtc_set_params": (C.c_int, [_P, C.POINTER(Params)]),
    "rtc_trace_closest": (C.c_int, [_P, C.c_int64, _P, _P, _P]),
    "rtc_camera_rays": (C.c_int, [_P, C.c_int64, _P, _P, _P]),
    "rtc_build_bvh_device": (C.c_int, [_P, C.c_int32, _P]),
    "rtc_debug_create_horizon": (C.c_int, [_P, C.c_int64, _P, _P]),
    "rtc_render": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, C.c_uint32]),
    "rtc_sync": (C.c_int, [_P]),
    "rtc_clear_accum": (C.c_int, [_P]),
    "rtc_read_accum": (C.c_int, [_P, _P, _P, _P]),
    "rtc_render_read": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, _P, _P]),
    "rtc_write_accum": (C.c_int, [_P, _P, _P, _P]),
    "rtc_accum_device_ptrs": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "rtc_tonemap_argb": (C.c_int, [_P, C.c_double, C.POINTER(C.c_double), C.c_double, _P]),
    "rtc_read_pixel": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "rtc_debug_trace": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_uint32, C.c_int32, C.POINTER(DebugRay), C.POINTER(C.c_int32)]),
    "rtc_debug_raycast": (C.c_int, [_P, C.c_int32, _P]),
    "rtc_debug_raycast_selection": (C.c_int, [_P, C.c_int32, _P, _P]),
    "rtc_render_samples": (C.c_int, [_P, C.c_uint32, _P]),
    "rtc_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "rtc_reset_stats": (C.c_int, [_P]),
    "rtc_comm_unique_id": (C.c_int, [_P]),
    "rtc_comm_init": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "rtc_reduce_accum": (C.c_int, [_P, C.c_int32]),
    "rtc_bcast_scene": (C.c_int, [_P, C.c_int32]),
    "rtc_comm_destroy": (C.c_int, [_P]),
    # rtcore_host.h
    "rtcs_scene_load": (_P, [C.c_char_p, C.c_char_p, C.c_int32]),
    "rtcs_scene_parse": (_P, [C.c_char_p, C.c_char_p, C.c_int32]),
    "rtcs_scene_synthetic": (_P, [C.c_char_p, C.c_int64, C.c_uint64, C.c_double]),
    "rtcs_scene_free": (None, [_P]),
    "rtcs_scene_globals": (C.c_int, [_P, C.POINTER(Globals)]),
    "rtcs_scene_override": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "rtcs_scene_set_ambient": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "rtcs_scene_set_debug_geom": (C.c_int, [_P, C.c_int32]),
    "rtcs_scene_desc": (C.c_int, [_P, C.POINTER(SceneDesc)]),
    "rtcs_scene_params": (C.c_int, [_P, C.c_uint64, C.POINTER(Params)]),
    "rtcs_scene_camera": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(Camera)]),
    "rtcs_scene_bvh": (C.c_int, [_P, C.POINTER(C.POINTER(BvhNode)), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "rtcs_scene_primitive_bounds": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "rtcs_desc_primitive_bounds": (C.c_int, [C.POINTER(SceneDesc), C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "rtcs_build_bvh": (C.c_int, [C.POINTER(SceneDesc), C.c_int32, C.POINTER(BvhNode), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "rtcs_raytracer_create": (_P, [_P, C.c_int32, C.c_int32, C.c_uint64, STATUS_FN, _P, C.c_char_p, C.c_int32]),
    "rtcs_raytracer_destroy": (None, [_P]),
    "rtcs_raytracer_start": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "rtcs_raytracer_stop": (None, [_P]),
    "rtcs_raytracer_pause": (None, [_P]),
    "rtcs_raytracer_resume": (None, [_P]),
    "rtcs_raytracer_is_running": (C.c_int, [_P]),
    "rtcs_raytracer_is_paused": (C.c_int, [_P]),
    "rtcs_raytracer_is_stopping": (C.c_int, [_P]),
    "rtcs_raytracer_set_exposure": (None, [_P, C.c_double]),
    "rtcs_raytracer_get_sample_set": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "rtcs_raytracer_get_bitmap": (C.c_int, [_P, _P]),
    "rtcs_raytracer_ctx": (_P, [_P]),
    "rtcs_raytracer_last_error": (C.c_char_p, [_P]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "librtcore_b200.so is not built (%s). Run `python -m raytracercore_b200.build`; "
            "there is no fallback implementation." % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        if HOST_ONLY and not (name.startswith("rtcs_scene_") or name == "rtcs_build_bvh"):
            continue
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class RtcError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("rtcore_b200 error %d: %s" % (code, message))
        self.code = code


def check(ctx, rc):
    if rc != RTC_OK:
        msg = lib.rtc_last_error(ctx)
        raise RtcError(rc, msg.decode() if msg else "")
