"""raytracercore_b200 — B200-native (sm_100a) wavefront path-tracing backend for Zaggy1024/RaytracerCore's render
loop. The product is librtcore_b200.so (C ABI in include/); this package is the thin Python view of it.

Attributes are resolved lazily so that `raytracercore_b200.build` can run before the shared library exists; any
other use loads the library and fails loudly if it is missing or lacks a declared symbol."""
import importlib

_EXPORTS = {
    "Scene": "scene", "LoaderException": "scene",
    "Context": "renderer", "FullRaytracer": "renderer", "RAY_DT": "renderer", "HIT_DT": "renderer",
    "RtcError": "_native", "RTC_F32": "_native", "RTC_F64": "_native", "RTC_OPT_COUNTERS": "_native",
    "RTC_OPT_KERNEL_TIMING": "_native", "RTC_OPT_MAX_PATHS": "_native", "RTC_OPT_WAVES": "_native", "RTC_OPT_REORDER": "_native",
    "RTC_BUILDER_SAH": "_native", "RTC_BUILDER_PLOC": "_native",
}
__all__ = sorted(_EXPORTS)


def __getattr__(name):
    if name in _EXPORTS:
        return getattr(importlib.import_module("." + _EXPORTS[name], __name__), name)
    if name in ("_native", "scene", "renderer", "build"):
        return importlib.import_module("." + name, __name__)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
