"""The C-ABI library loads and exports every symbol include/*.h declares; without a GPU it fails loudly."""
import ctypes as C
import os
import re

import pytest

from raytracercore_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in ("rtcore_b200.h", "rtcore_host.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"\b(rtcs?_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_every_declared_symbol_is_exported():
    syms = declared_symbols()
    assert len(syms) >= 45
    lib = C.CDLL(N.LIB_PATH)
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_binding_table_matches_headers():
    assert sorted(N.SIGNATURES) == [s for s in declared_symbols() if s != "rtcs_status_fn"]


def test_abi_version_and_struct_sizes():
    assert N.lib.rtc_abi_version() == 1
    assert C.sizeof(N.BvhNode) == 64 and C.sizeof(N.Ray) == 48 and C.sizeof(N.Hit) == 64
    assert C.sizeof(N.Camera) == 8 + 12 * 8 + 9 * 8
    assert C.sizeof(N.Params) == 16 + 24 + 8 + 8


def test_no_cpu_fallback():
    if N.lib.rtc_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = N.lib.rtc_create(0, N.RTC_F32, C.byref(h))
    assert rc == N.RTC_ERR_CUDA and not h.value
    assert b"no CPU fallback" in N.lib.rtc_last_error(None)


def test_argument_checking_without_device():
    assert N.lib.rtc_create(0, 7, C.byref(C.c_void_p())) == N.RTC_ERR_INVALID
    assert N.lib.rtc_set_option(None, 1, 1) == N.RTC_ERR_INVALID
    assert N.lib.rtc_render(None, 0, 0, 1, 1, 0, 1) == N.RTC_ERR_INVALID
