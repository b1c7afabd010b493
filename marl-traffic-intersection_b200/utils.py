"""Route tables the reference's env.py relies on (/root/reference/utils.py:29-52).

Only the two default route mappings matter to the simulation; ``build_lane_layout`` of the reference
(utils.py:55-98) is computed for a 900x900 canvas and never reaches the C++ core (env.py:108-109), so it
is not reproduced: lane points come from the library (``isx_route`` / RouteGen.cpp:7-53 restated)."""

OBS_DIM = 127

def _mapping(exits):
    """{"IN_k": ["OUT_e"]} for every entry lane k = 1.. that has a default exit e (None = no default route)."""
    return {f"IN_{k}": [f"OUT_{e}"] for k, e in enumerate(exits, start=1) if e is not None}


# default exit lane of IN_1, IN_2, ... (utils.py:29-52 of the reference); insertion order = entry-lane order, which is
# the order env.py iterates the mapping in (env.py:118-122, 138-145)
DEFAULT_ROUTE_MAPPING_2LANES = _mapping((3, 6, 5, 8, None, 2, 1, 4))
DEFAULT_ROUTE_MAPPING_3LANES = _mapping((4, 8, 12, 7, 11, 3, 10, 2, 6, 1, 5, 9))

# env.py:41-54: reward weights in reward_vector() order
_REWARD_KEYS = ("progress_scale", "stuck_speed_threshold", "stuck_penalty", "crash_vehicle_penalty", "crash_object_penalty",
                "success_reward", "action_smoothness_scale", "team_alpha")
DEFAULT_REWARD_CONFIG = {"use_team_reward": False, "traffic_flow": False,
                         "reward_config": dict(zip(_REWARD_KEYS, (10.0, 1.0, -0.01, -10.0, -5.0, 10.0, -0.02, 0.2)))}

STATUS_NAMES = ("ALIVE", "DEAD", "SUCCESS", "CRASH_WALL", "CRASH_LINE", "CRASH_CAR")


def all_default_routes(num_lanes: int):
    """(start, end) pairs in the dict-iteration order env.py uses (env.py:118-122, 138-145)."""
    mapping = DEFAULT_ROUTE_MAPPING_2LANES if num_lanes == 2 else DEFAULT_ROUTE_MAPPING_3LANES
    return [(s, e) for s, ends in mapping.items() for e in ends]


def default_ego_routes(num_agents: int, num_lanes: int):
    routes = all_default_routes(num_lanes)
    return [routes[i % len(routes)] for i in range(num_agents)]


def reward_vector(reward_cfg=None):
    """dict with env.py's keys (env.py:57-77) -> (k_prog, v_min_ms, k_stuck, k_cv, k_co, k_succ, k_sm, alpha)."""
    base = dict(DEFAULT_REWARD_CONFIG["reward_config"])
    if isinstance(reward_cfg, dict):
        base.update({k: v for k, v in reward_cfg.items() if k in base})
    return (
        float(base["progress_scale"]), float(base["stuck_speed_threshold"]), float(base["stuck_penalty"]),
        float(base["crash_vehicle_penalty"]), float(base["crash_object_penalty"]), float(base["success_reward"]),
        float(base["action_smoothness_scale"]), float(base["team_alpha"]),
    )
