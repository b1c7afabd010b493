"""Module-level names of the reference's utils.py (/root/reference/utils.py): canvas constants, the two default route
mappings (utils.py:29-52), the default reward weights and ``build_lane_layout`` (utils.py:55-98).

Only the route mappings matter to the simulation: ``build_lane_layout`` describes a 900x900 canvas that never reaches the
C++ core (env.py:108-109 only stores it); the 750-px lane points the simulation uses come from the library
(``isx_route`` / RouteGen.cpp:7-53 restated)."""

WIDTH, HEIGHT = 900, 900          # utils.py:4 (the Python-side canvas; the simulation's own is 750 px, constants.h:4-5)
SCALE = 12
LANE_WIDTH_M = 3.5
LANE_WIDTH_PX = int(LANE_WIDTH_M * SCALE)

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


def build_lane_layout(num_lanes: int):
    """Lane end points of an N/E/S/W intersection on the WIDTH x HEIGHT canvas, 30 px in from the edge: lane j of a
    direction sits LANE_WIDTH_PX * (j + 0.5) from the centre line, entry lanes on the right-hand side of travel
    (utils.py:55-98).  Ids count N, E, S, W x lane: IN_1.. / OUT_1.. ."""
    order = ["N", "E", "S", "W"]
    cx, cy, margin = WIDTH // 2, HEIGHT // 2, 30
    # per direction: fixed edge coordinate, which axis varies, sign of the entry-lane offset
    edge = {"N": ("y", margin, -1), "S": ("y", HEIGHT - margin, +1), "E": ("x", WIDTH - margin, -1), "W": ("x", margin, +1)}
    layout = {"points": {}, "in_by_dir": {d: [] for d in order}, "out_by_dir": {d: [] for d in order}, "dir_of": {}, "idx_of": {},
              "dir_order": order}
    for di, d in enumerate(order):
        fixed_axis, fixed, sign = edge[d]
        for j in range(num_lanes):
            off = LANE_WIDTH_PX * (0.5 + j)
            for kind, sg in (("IN", sign), ("OUT", -sign)):
                name = f"{kind}_{di * num_lanes + j + 1}"
                moving = (cx if fixed_axis == "y" else cy) + sg * off
                layout["points"][name] = (moving, fixed) if fixed_axis == "y" else (fixed, moving)
                layout["in_by_dir" if kind == "IN" else "out_by_dir"][d].append(name)
                layout["dir_of"][name], layout["idx_of"][name] = d, j
    return layout


def reward_vector(reward_cfg=None):
    """dict with env.py's keys (env.py:57-77), or an 8-sequence -> (k_prog, v_min_ms, k_stuck, k_cv, k_co, k_succ, k_sm, alpha)."""
    if reward_cfg is not None and not isinstance(reward_cfg, dict):
        vec = tuple(float(x) for x in reward_cfg)
        if len(vec) != 8:
            raise ValueError("reward vector needs 8 entries")
        return vec
    base = dict(DEFAULT_REWARD_CONFIG["reward_config"])
    if isinstance(reward_cfg, dict):
        base.update({k: v for k, v in reward_cfg.items() if k in base})
    return (
        float(base["progress_scale"]), float(base["stuck_speed_threshold"]), float(base["stuck_penalty"]),
        float(base["crash_vehicle_penalty"]), float(base["crash_object_penalty"]), float(base["success_reward"]),
        float(base["action_smoothness_scale"]), float(base["team_alpha"]),
    )
