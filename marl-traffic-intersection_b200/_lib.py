"""ctypes binding of libisx_b200.so (include/isx.h).  The library is built in-tree by build.py; there is
no CPU fallback: a missing library or a machine without a CUDA device is an error, loudly."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ISX_LIB") or os.path.join(HERE, "csrc", "libisx_b200.so")   # ISX_LIB: tuning builds (tools/)

ISX_ABI_VERSION = 2
OBS_DIM = 127
MAX_RAYS = 96
STATS_COUNTERS = 15          # isx.h ISX_STATS_COUNTERS; indices ISX_STAT_*
STAT_INDEX = {"npc_spawned": 6, "npc_removed": 7, "npc_collided": 8, "npc_overflow": 9,
              "env_resets": 10, "agent_steps": 11, "neighbor_tie_sorts": 12}
E_ARG, E_CUDA, E_ROUTE_START, E_ROUTE_END, E_STATE = -1, -2, -3, -4, -5


class IsxError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"isx error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("num_envs", C.c_int32), ("num_agents", C.c_int32),
        ("num_lanes", C.c_int32), ("lidar_rays", C.c_int32), ("npc_capacity", C.c_int32),
        ("use_team_reward", C.c_int32), ("respawn_enabled", C.c_int32), ("max_steps", C.c_int32),
        ("traffic_flow", C.c_int32), ("traffic_density", C.c_float), ("reward", C.c_float * 8),
        ("ego_start", C.POINTER(C.c_char_p)), ("ego_end", C.POINTER(C.c_char_p)),
        ("num_traffic_routes", C.c_int32),
        ("traffic_start", C.POINTER(C.c_char_p)), ("traffic_end", C.POINTER(C.c_char_p)),
        ("seed", C.c_uint64), ("env_id_base", C.c_int64), ("auto_reset", C.c_int32), ("reserved", C.c_int32),
    ]


class CarState(C.Structure):
    _fields_ = [
        ("x", C.c_float), ("y", C.c_float), ("v", C.c_float), ("heading", C.c_float),
        ("acc", C.c_float), ("steer", C.c_float),
        ("prev_dist", C.c_float), ("prev_a0", C.c_float), ("prev_a1", C.c_float),
        ("path_index", C.c_int32), ("route", C.c_int32), ("alive", C.c_int32), ("uid", C.c_uint32), ("intention", C.c_int32),
    ]


class TrafficEvents(C.Structure):
    _fields_ = [("rng_draws", C.c_int32), ("spawn_route", C.c_int32), ("spawned", C.c_int32),
                ("removed_mask", C.c_uint32), ("collided_mask", C.c_uint32), ("npc_count", C.c_int32)]


_BUF_FIELDS = [
    ("obs", "f4"), ("reward", "f4"), ("done", "u1"), ("status", "u1"), ("terminated", "u1"), ("truncated", "u1"),
    ("agents_alive", "i4"), ("step", "i4"), ("lidar_hit", "u1"),
    ("ego_x", "f4"), ("ego_y", "f4"), ("ego_v", "f4"), ("ego_heading", "f4"), ("ego_steer", "f4"), ("ego_acc", "f4"),
    ("ego_prev_dist", "f4"), ("ego_prev_a0", "f4"), ("ego_prev_a1", "f4"), ("ego_path_index", "i4"), ("ego_alive", "u1"),
    ("npc_x", "f4"), ("npc_y", "f4"), ("npc_v", "f4"), ("npc_heading", "f4"), ("npc_steer", "f4"),
    ("npc_path_index", "i4"), ("npc_route", "i4"), ("npc_uid", "u4"), ("npc_count", "i4"),
    ("events", "V24"), ("tick", "u4"),
]


class Buffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n, _ in _BUF_FIELDS]


class Stats(C.Structure):
    _fields_ = [("agent_steps", C.c_int64), ("status_hist", C.c_int64 * 6), ("npc_spawned", C.c_int64),
                ("npc_removed", C.c_int64), ("npc_collided", C.c_int64), ("npc_overflow", C.c_int64),
                ("env_resets", C.c_int64), ("reward_sum", C.c_double), ("neighbor_tie_sorts", C.c_int64)]


EXPORTS = [
    "isx_last_error", "isx_abi_version", "isx_create", "isx_create_groups", "isx_num_groups", "isx_group_range", "isx_destroy", "isx_reset", "isx_step", "isx_step_host",
    "isx_step_pinned", "isx_host_views", "isx_host_views_aux", "isx_host_step_info", "isx_expand_obs_rows",
    "isx_rollout", "isx_rollout_timed", "isx_rollout_timed4", "isx_get_buffers", "isx_num_envs", "isx_num_agents", "isx_get_env_state", "isx_set_env_state",
    "isx_observe", "isx_render", "isx_snapshot_create", "isx_snapshot_save", "isx_snapshot_restore", "isx_snapshot_destroy",
    "isx_stats_read", "isx_stats_reset", "isx_trace_read", "isx_debug_check_guards", "isx_pipe_timeline", "isx_stats_device_ptrs", "isx_route", "isx_math_probe",
    "isx_set_lidar_rays", "isx_lidar_rays", "isx_set_reward", "isx_configure_episode", "isx_set_traffic_density",
    "isx_car_update", "isx_car_check_collision",
]

_lib = None


def load_library(path: str | None = None):
    """Load libisx_b200.so and declare every prototype of include/isx.h.  Raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"libisx_b200.so not found at {p}: build it with `python -m marl_traffic_intersection_b200.build` "
            "(or __graft_entry__.build()).  There is no CPU fallback."
        )
    lib = C.CDLL(p)
    vp, i32, f32 = C.c_void_p, C.c_int32, C.c_float
    lib.isx_last_error.restype = C.c_char_p
    lib.isx_last_error.argtypes = []
    lib.isx_abi_version.restype = C.c_int
    lib.isx_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.isx_create_groups.argtypes = [C.POINTER(Config), i32, C.POINTER(vp)]
    lib.isx_num_groups.argtypes = [vp]
    lib.isx_group_range.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32)]
    lib.isx_destroy.argtypes = [vp]
    lib.isx_reset.argtypes = [vp, vp, vp]
    lib.isx_step.argtypes = [vp, vp, f32, vp]
    lib.isx_step_host.argtypes = [vp, vp, f32, vp, vp, vp, vp, vp, vp, vp]
    lib.isx_step_pinned.argtypes = [vp, f32, vp]
    lib.isx_host_views.argtypes = [vp] + [C.POINTER(vp)] * 7
    lib.isx_host_views_aux.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    lib.isx_expand_obs_rows.argtypes = [vp, vp, i32, vp, C.c_int64]
    lib.isx_host_step_info.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(i32), C.POINTER(i32)]
    lib.isx_rollout.argtypes = [vp, i32, f32, vp]
    lib.isx_rollout_timed.argtypes = [vp, i32, f32, vp, C.POINTER(f32), C.POINTER(f32)]
    lib.isx_rollout_timed4.argtypes = [vp, i32, f32, vp, C.POINTER(f32)]
    lib.isx_get_buffers.argtypes = [vp, C.POINTER(Buffers)]
    lib.isx_num_envs.argtypes = [vp]
    lib.isx_num_agents.argtypes = [vp]
    lib.isx_get_env_state.argtypes = [vp, i32, C.POINTER(CarState), C.POINTER(CarState), i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_uint32)]
    lib.isx_set_env_state.argtypes = [vp, i32, C.POINTER(CarState), C.POINTER(CarState), i32, i32, C.c_uint32]
    lib.isx_observe.argtypes = [vp, vp]
    lib.isx_render.argtypes = [vp, i32, vp, vp]
    lib.isx_snapshot_create.argtypes = [vp, C.POINTER(vp)]
    lib.isx_snapshot_save.argtypes = [vp, vp, vp]
    lib.isx_snapshot_restore.argtypes = [vp, vp, vp, vp]
    lib.isx_snapshot_destroy.argtypes = [vp]
    lib.isx_stats_read.argtypes = [vp, C.POINTER(Stats)]
    lib.isx_stats_reset.argtypes = [vp]
    lib.isx_trace_read.argtypes = [vp, vp]
    lib.isx_debug_check_guards.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.isx_pipe_timeline.argtypes = [vp, f32, vp, C.POINTER(f32), i32]
    lib.isx_stats_device_ptrs.argtypes = [vp, C.POINTER(vp), C.POINTER(i32), C.POINTER(vp), vp]
    lib.isx_set_lidar_rays.argtypes = [vp, i32]
    lib.isx_lidar_rays.argtypes = [vp]
    lib.isx_set_reward.argtypes = [vp, i32, C.POINTER(f32)]
    lib.isx_configure_episode.argtypes = [vp, i32, i32, i32, i32]
    lib.isx_set_traffic_density.argtypes = [vp, i32, f32]
    lib.isx_car_update.argtypes = [i32, C.POINTER(CarState), f32, f32, f32]
    lib.isx_car_check_collision.argtypes = [i32, C.POINTER(CarState), C.POINTER(CarState), C.POINTER(i32)]
    lib.isx_route.argtypes = [i32, C.c_char_p, C.c_char_p, vp, C.POINTER(i32), C.POINTER(f32), C.POINTER(f32), C.POINTER(f32)]
    lib.isx_math_probe.argtypes = [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    for n in EXPORTS:
        if n not in ("isx_last_error",):
            getattr(lib, n).restype = C.c_int
    if lib.isx_abi_version() != ISX_ABI_VERSION:
        raise RuntimeError("libisx_b200.so ABI version mismatch")
    if path is None:
        _lib = lib
    return lib


def check(lib, rc: int) -> int:
    if rc < 0:
        raise IsxError(rc, (lib.isx_last_error() or b"").decode())
    return rc
