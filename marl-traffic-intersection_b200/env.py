"""Single-env facade with the reference's Python API (/root/reference/env.py:80-221), backed by the B200
stepper.  Same constructor config keys, same reset()/step() return shapes and dtypes, same info dict keys,
same exceptions at the same points.  For throughput use BatchedIntersectionEnv; this class exists so that
code written against the reference's env.py runs unchanged (one env on the GPU is latency-, not
throughput-oriented)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import cpp_backend
from .utils import DEFAULT_REWARD_CONFIG, all_default_routes, build_lane_layout

# reward_config key -> RewardConfig attribute (env.py:57-77)
_REWARD_ATTR = {
    "progress_scale": "k_prog", "stuck_speed_threshold": "v_min_ms", "stuck_penalty": "k_stuck",
    "crash_vehicle_penalty": "k_cv", "crash_object_penalty": "k_co", "success_reward": "k_succ",
    "action_smoothness_scale": "k_sm", "team_alpha": "alpha",
}


class IntersectionEnv:
    """One intersection env.  With ``traffic_flow`` the reference forces a single ego and returns flat (127,) obs and a
    float reward (env.py:87-90,157-161,205-207); otherwise obs is (N,127) and rewards (N,)."""

    def __init__(self, config: Optional[Dict[str, Any]] = None):
        opt = dict(config or {})
        get = opt.get
        self.traffic_flow = bool(get("traffic_flow", False))
        single = self.traffic_flow                                    # NPC traffic => exactly one ego, no team reward
        self.num_agents = 1 if single else int(get("num_agents", 1))
        self.num_lanes = int(get("num_lanes", 3))
        self.render_mode = get("render_mode", None)
        self.show_lane_ids, self.show_lidar = bool(get("show_lane_ids", False)), bool(get("show_lidar", False))
        self.traffic_density = float(get("traffic_density", 0.5))
        table = all_default_routes(self.num_lanes)                    # dict order of the default mapping (env.py:118-123)
        routes = get("ego_routes", None)
        self.ego_routes: Sequence[Tuple[str, str]] = routes if routes is not None else [table[i % len(table)] for i in range(self.num_agents)]

        self.lane_layout = build_lane_layout(self.num_lanes)           # env.py:108-109 (Python-side only)
        self.points = self.lane_layout["points"]

        core = cpp_backend.IntersectionEnv(self.num_lanes)
        core.seed, core.lidar_rays = int(get("seed", 0)), int(get("lidar_rays", 96))
        team = False if single else bool(get("use_team_reward", DEFAULT_REWARD_CONFIG.get("use_team_reward", False)))
        core.configure(team, bool(get("respawn_enabled", True)), int(get("max_steps", 2000)))
        core.configure_traffic(self.traffic_flow, self.traffic_density)
        core.configure_routes(table)
        weights = get("reward_config", None)
        if weights is None:
            weights = DEFAULT_REWARD_CONFIG.get("reward_config", {})
        if isinstance(weights, dict):
            for key, attr in _REWARD_ATTR.items():
                if key in weights:
                    setattr(core.reward_config, attr, float(weights[key]))
        self.env = core
        self._cars: Optional[List["cpp_backend.Car"]] = None           # env.py:152,155,184 snapshot cars / traffic_cars eagerly;
        self._traffic_cars: Optional[List["cpp_backend.Car"]] = None   # here the device round trip is paid only if somebody looks
        self.reset()

    # ------------------------------------------------------------------ state views
    @property
    def cars(self) -> List["cpp_backend.Car"]:
        """env.cars as of the last reset() (the reference keeps that list object, env.py:152)."""
        return self._cars

    @property
    def traffic_cars(self) -> List["cpp_backend.Car"]:
        """NPC snapshot after the last reset()/step() (env.py:155,184); empty without traffic_flow."""
        if not self.traffic_flow:
            return []
        if self._traffic_cars is None:
            self._traffic_cars = list(self.env.traffic_cars)
        return self._traffic_cars

    def _obs(self, rows) -> np.ndarray:
        arr = np.asarray(rows, dtype=np.float32)
        return arr[0] if self.traffic_flow else arr

    # ------------------------------------------------------------------ gym-style surface
    def reset(self):
        core = self.env
        core.reset()
        for start_id, end_id in list(self.ego_routes)[: self.num_agents]:
            core.add_car_with_route(start_id, end_id)
        if len(self.ego_routes) < self.num_agents:                    # same failure as indexing past the list in env.py:150
            raise IndexError("list index out of range")
        self._cars = core.cars                                        # reset-time snapshot (resets are rare)
        self._traffic_cars = None
        return self._obs(core.get_observations()), {}

    def step(self, actions, dt: float = 1.0 / 60.0):
        act = np.asarray(actions, dtype=np.float32)
        if self.traffic_flow or (act.ndim == 1 and act.size == 2 and self.num_agents == 1):
            act = act.reshape(1, 2)
        elif act.ndim == 1:
            raise ValueError(f"Expected actions shape (N,2) for multi-agent, got {act.shape}")
        out = self.env.step(act[:, 0].tolist(), act[:, 1].tolist(), float(dt))
        self._traffic_cars = None                                     # refetched on access (state cannot change between steps)
        rew = np.asarray(out.rewards, dtype=np.float32)
        scalar = float(rew[0]) if rew.size else 0.0
        done_flags = (bool(out.terminated), bool(out.truncated))
        info = dict(step=int(out.step), rewards=scalar if self.traffic_flow else rew.tolist(),
                    collisions={int(a): str(s) for a, s in zip(out.agent_ids, out.status)},
                    agents_alive=int(out.agents_alive), terminated=done_flags[0], truncated=done_flags[1],
                    done=list(out.done), status=list(out.status))
        return self._obs(out.obs), (scalar if self.traffic_flow else rew), done_flags[0], done_flags[1], info

    def render(self, show_lane_ids: Optional[bool] = None, show_lidar: Optional[bool] = None):
        """env.py:210-217: nothing unless render_mode == "human".  The reference then draws into its Windows/GLFW window;
        here the same call returns the headless picture (uint8 [750, 750, 3]) of isx_render."""
        if self.render_mode != "human":
            return None
        lane_ids = self.show_lane_ids if show_lane_ids is None else show_lane_ids
        lidar = self.show_lidar if show_lidar is None else show_lidar
        return self.env.render(bool(lane_ids), bool(lidar))

    def close(self):
        if getattr(self.env, "_benv", None) is not None:
            self.env._benv.close()
            self.env._benv = None
