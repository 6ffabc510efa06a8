"""raytracercore_b200 — B200-native (sm_100a) wavefront path-tracing backend for Zaggy1024/RaytracerCore's render
loop. The product is librtcore_b200.so (C ABI in include/); this package is the thin Python view of it."""
from . import _native
from ._native import (RTC_F32, RTC_F64, RTC_OPT_COUNTERS, RTC_OPT_KERNEL_TIMING, RTC_OPT_MAX_PATHS, RtcError)
from .renderer import HIT_DT, RAY_DT, Context, FullRaytracer
from .scene import LoaderException, Scene

__all__ = ["Scene", "Context", "FullRaytracer", "LoaderException", "RtcError", "RTC_F32", "RTC_F64", "RAY_DT", "HIT_DT",
           "RTC_OPT_COUNTERS", "RTC_OPT_KERNEL_TIMING", "RTC_OPT_MAX_PATHS", "_native"]
