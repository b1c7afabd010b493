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
    cfg.abi_version, cfg.num_envs, cfg.num_agents, cfg.num_lanes, cfg.lidar_rays = _lib.ISX_ABI_VERSION, 1, 1, 3, 96
    s, e = (C.c_char_p * 1)(b"IN_6"), (C.c_char_p * 1)(b"OUT_2")
    cfg.ego_start, cfg.ego_end = s, e
    h = C.c_void_p()
    assert lib.isx_create(C.byref(cfg), C.byref(h)) == _lib.E_CUDA
    assert b"no CUDA device" in lib.isx_last_error()


def test_config_validation_happens_before_any_device_work():
    """Argument checks of isx_create / isx_create_groups need no GPU: bad sizes, bad ABI version, groups that disagree on
    what shapes the shared buffers; a well-formed request then fails with ISX_E_CUDA on a box without a device."""
    import torch
    from marl_traffic_intersection_b200 import _lib
    lib = _lib.load_library()
    s, e = (C.c_char_p * 3)(b"IN_6", b"IN_4", b"IN_5"), (C.c_char_p * 3)(b"OUT_2", b"OUT_8", b"OUT_7")

    def cfg(**kw):
        c = _lib.Config()
        c.abi_version, c.num_envs, c.num_agents, c.num_lanes, c.lidar_rays = _lib.ISX_ABI_VERSION, 4, 3, 3, 96
        c.ego_start, c.ego_end = s, e
        for k, v in kw.items():
            setattr(c, k, v)
        return c

    h = C.c_void_p()
    for bad in (cfg(abi_version=99), cfg(num_envs=0), cfg(num_agents=0), cfg(num_agents=33), cfg(num_lanes=5), cfg(lidar_rays=0),
                cfg(lidar_rays=97), cfg(npc_capacity=33), cfg(num_traffic_routes=-1)):
        assert lib.isx_create(C.byref(bad), C.byref(h)) == _lib.E_ARG and not h.value
    assert lib.isx_create(None, C.byref(h)) == _lib.E_ARG
    two = (_lib.Config * 2)(cfg(), cfg(num_agents=2))
    assert lib.isx_create_groups(two, 2, C.byref(h)) == _lib.E_ARG and b"same in every group" in lib.isx_last_error()
    two = (_lib.Config * 2)(cfg(), cfg(lidar_rays=72))
    assert lib.isx_create_groups(two, 2, C.byref(h)) == _lib.E_ARG
    two = (_lib.Config * 2)(cfg(traffic_flow=1, npc_capacity=16), cfg(traffic_flow=1, npc_capacity=8))
    assert lib.isx_create_groups(two, 2, C.byref(h)) == _lib.E_ARG and b"npc_capacity" in lib.isx_last_error()
    assert lib.isx_create_groups(two, 0, C.byref(h)) == _lib.E_ARG and lib.isx_create_groups(two, 65, C.byref(h)) == _lib.E_ARG
    if not torch.cuda.is_available():
        ok = (_lib.Config * 2)(cfg(), cfg(num_lanes=2))
        assert lib.isx_create_groups(ok, 2, C.byref(h)) == _lib.E_CUDA


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "marl-traffic-intersection_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "pyoracle" not in txt and "isx_oracle" not in txt and "libisx_ref" not in txt, f
