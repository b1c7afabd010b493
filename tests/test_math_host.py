"""Restated libm (csrc/isx_math.cuh, host build) against this machine's glibc — the entry points the reference binds.
The three one-argument functions are swept over ALL 2^32 float bit patterns; two-argument ones over structured and
random pairs.  ~2 minutes on 8 cores."""
import ctypes as C
import os
import struct

import pytest

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "marl-traffic-intersection_b200", "csrc", "libisx_math_host.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="libisx_math_host.so not built")
NT = os.cpu_count() or 4


def lib():
    l = C.CDLL(LIB)
    for n in ("isxm_sweep_sincosf", "isxm_sweep_sinf_cosf", "isxm_sweep_tanf", "isxm_sweep_atanf", "isxm_sweep_fmod"):
        f = getattr(l, n)
        f.restype = C.c_uint64
        f.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_uint32)]
    l.isxm_sweep_pair.restype = C.c_uint64
    l.isxm_sweep_pair.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_uint32)]
    l.isxm_random_pairs.restype = C.c_uint64
    l.isxm_random_pairs.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    return l


@pytest.mark.parametrize("fn", ["isxm_sweep_sincosf", "isxm_sweep_sinf_cosf", "isxm_sweep_tanf", "isxm_sweep_atanf", "isxm_sweep_fmod"])
def test_exhaustive_one_argument(fn):
    fb = C.c_uint32()
    # fmod: all |x| < 2^20 of both signs + inf/nan patterns (libm's generic fmodf is slow on huge quotients)
    ranges = [(0, 0x49800000), (0x7F000000, 0x7FFFFFFF), (0x80000000, 0xC9800000), (0xFF000000, 0xFFFFFFFF)] if fn == "isxm_sweep_fmod" else [(0, 0xFFFFFFFF)]
    for lo, hi in ranges:
        bad = getattr(lib(), fn)(lo, hi, NT, C.byref(fb))
        assert bad == 0, f"{fn}: {bad} mismatches, first at bits {fb.value:#x}"


@pytest.mark.parametrize("other", [1.0, -0.0, -270.0])
def test_pair_sweeps(other):
    ob = struct.unpack("<I", struct.pack("<f", other))[0]
    fb = C.c_uint32()
    # all finite |x| < 2^20 of both signs against a fixed partner: atan2f(x,o), atan2f(o,x), hypotf(x,o)
    for lo, hi in ((0, 0x49800000), (0x80000000, 0xC9800000)):
        bad = lib().isxm_sweep_pair(lo, hi, ob, NT, C.byref(fb))
        assert bad == 0, f"pair sweep vs {other}: {bad} mismatches, first at {fb.value:#x}"


@pytest.mark.parametrize("mode", [0, 1])
def test_random_pairs(mode):
    by, bx = C.c_float(), C.c_float()
    bad = lib().isxm_random_pairs(99, 200_000_000, mode, NT, C.byref(by), C.byref(bx))
    assert bad == 0, (by.value, bx.value)
