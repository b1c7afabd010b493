"""Route tables the reference's env.py relies on (/root/reference/utils.py:29-52).

Only the two default route mappings matter to the simulation; ``build_lane_layout`` of the reference
(utils.py:55-98) is computed for a 900x900 canvas and never reaches the C++ core (env.py:108-109), so it
is not reproduced: lane points come from the library (``isx_route`` / RouteGen.cpp:7-53 restated)."""

OBS_DIM = 127

DEFAULT_ROUTE_MAPPING_2LANES = {
    "IN_1": ["OUT_3"],
    "IN_2": ["OUT_6"],
    "IN_3": ["OUT_5"],
    "IN_4": ["OUT_8"],
    "IN_6": ["OUT_2"],
    "IN_7": ["OUT_1"],
    "IN_8": ["OUT_4"],
}

DEFAULT_ROUTE_MAPPING_3LANES = {
    "IN_1": ["OUT_4"],
    "IN_2": ["OUT_8"],
    "IN_3": ["OUT_12"],
    "IN_4": ["OUT_7"],
    "IN_5": ["OUT_11"],
    "IN_6": ["OUT_3"],
    "IN_7": ["OUT_10"],
    "IN_8": ["OUT_2"],
    "IN_9": ["OUT_6"],
    "IN_10": ["OUT_1"],
    "IN_11": ["OUT_5"],
    "IN_12": ["OUT_9"],
}

DEFAULT_REWARD_CONFIG = {
    "use_team_reward": False,
    "traffic_flow": False,
    "reward_config": {
        "progress_scale": 10.0,
        "stuck_speed_threshold": 1.0,
        "stuck_penalty": -0.01,
        "crash_vehicle_penalty": -10.0,
        "crash_object_penalty": -5.0,
        "success_reward": 10.0,
        "action_smoothness_scale": -0.02,
        "team_alpha": 0.2,
    },
}

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
