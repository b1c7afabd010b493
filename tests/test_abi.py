"""The C-ABI library loads on a machine without a GPU and exports every symbol include/isx.h declares; the Python
mirror of the reference interface raises the reference's exception types; no compute is attempted here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "isx.h")
LIB = os.path.join(ROOT, "marl-traffic-intersection_b200", "csrc", "libisx_b200.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="libisx_b200.so not built (run __graft_entry__.build())")


def declared():
    txt = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(isx_[a-z_0-9]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported():
    lib = C.CDLL(LIB)
    names = declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/isx.h but not exported"


def test_python_binding_lists_the_same_exports():
    from marl_traffic_intersection_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared()
    lib = _lib.load_library()
    assert lib.isx_abi_version() == _lib.ISX_ABI_VERSION


def test_struct_layouts_match_the_header():
    from marl_traffic_intersection_b200 import _lib
    assert C.sizeof(_lib.CarState) == 56 and C.sizeof(_lib.TrafficEvents) == 24
    assert C.sizeof(_lib.Stats) == 8 * 14
    assert C.sizeof(_lib.Buffers) == 8 * len(_lib._BUF_FIELDS)
    assert _lib.Config.seed.offset % 8 == 0 and C.sizeof(_lib.Config) == _lib.Config.reserved.offset + 4


def test_route_probe_and_error_codes_without_gpu():
    from marl_traffic_intersection_b200 import _lib
    lib = _lib.load_library()
    buf = (C.c_float * 320)()
    intent = C.c_int32()
    sx, sy, sh = C.c_float(), C.c_float(), C.c_float()
    n = lib.isx_route(3, b"IN_6", b"OUT_2", buf, C.byref(intent), C.byref(sx), C.byref(sy), C.byref(sh))
    assert n == 160 and intent.value == 2 and (sx.value, sy.value) == (720.0, 270.0)
    assert lib.isx_route(3, b"IN_99", b"OUT_2", None, None, None, None, None) == _lib.E_ROUTE_START
    assert lib.isx_route(3, b"IN_6", b"OUT_99", None, None, None, None, None) == _lib.E_ROUTE_END
    assert b"OUT_99" in lib.isx_last_error()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from marl_traffic_intersection_b200 import BatchedIntersectionEnv, IntersectionEnv, _lib
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        BatchedIntersectionEnv({"num_envs": 2})
    with pytest.raises(RuntimeError):
        IntersectionEnv({"num_agents": 1})
    # the C entry point itself refuses without a device
    lib = _lib.load_library()
    cfg = _lib.Config()
    cfg.abi_version, cfg.num_envs, cfg.num_agents, cfg.num_lanes, cfg.lidar_rays = 1, 1, 1, 3, 96
    s, e = (C.c_char_p * 1)(b"IN_6"), (C.c_char_p * 1)(b"OUT_2")
    cfg.ego_start, cfg.ego_end = s, e
    h = C.c_void_p()
    assert lib.isx_create(C.byref(cfg), C.byref(h)) == _lib.E_CUDA
    assert b"no CUDA device" in lib.isx_last_error()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "marl-traffic-intersection_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "pyoracle" not in txt and "isx_oracle" not in txt and "libisx_ref" not in txt, f
