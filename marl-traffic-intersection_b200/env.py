"""Single-env facade with the reference's Python API (/root/reference/env.py:80-221), backed by the B200
stepper.  Same constructor config keys, same reset()/step() return shapes and dtypes, same info dict keys,
same exceptions at the same points.  For throughput use BatchedIntersectionEnv; this class exists so that
code written against the reference's env.py runs unchanged (one env on the GPU is latency-, not
throughput-oriented)."""
from __future__ import annotations

from typing import Any, Dict, List, Union

import numpy as np

from . import cpp_backend
from .utils import DEFAULT_REWARD_CONFIG, DEFAULT_ROUTE_MAPPING_2LANES, DEFAULT_ROUTE_MAPPING_3LANES


def _apply_reward_config(env: Any, reward_cfg: Dict[str, Any]) -> None:   # env.py:57-77
    rc = env.reward_config
    for key, attr in (("progress_scale", "k_prog"), ("stuck_speed_threshold", "v_min_ms"), ("stuck_penalty", "k_stuck"),
                      ("crash_vehicle_penalty", "k_cv"), ("crash_object_penalty", "k_co"), ("success_reward", "k_succ"),
                      ("action_smoothness_scale", "k_sm"), ("team_alpha", "alpha")):
        if key in reward_cfg:
            setattr(rc, attr, float(reward_cfg[key]))


class IntersectionEnv:
    def __init__(self, config: Dict[str, Any] | None = None):
        if config is None:
            config = {}
        self.traffic_flow = bool(config.get("traffic_flow", False))
        self.num_agents = 1 if self.traffic_flow else int(config.get("num_agents", 1))      # env.py:87-90
        self.num_lanes = int(config.get("num_lanes", 3))
        self.render_mode = config.get("render_mode", None)
        self.show_lane_ids = bool(config.get("show_lane_ids", False))
        self.show_lidar = bool(config.get("show_lidar", False))
        use_team = bool(config.get("use_team_reward", DEFAULT_REWARD_CONFIG.get("use_team_reward", False)))
        if self.traffic_flow:
            use_team = False
        respawn = bool(config.get("respawn_enabled", True))
        max_steps = int(config.get("max_steps", 2000))
        self.ego_routes = config.get("ego_routes", None)
        if self.ego_routes is None:
            self.ego_routes = self._default_routes(self.num_agents, self.num_lanes)

        self.env = cpp_backend.IntersectionEnv(self.num_lanes)
        self.env.seed = int(config.get("seed", 0))
        self.env.lidar_rays = int(config.get("lidar_rays", 96))
        self.env.configure(use_team, respawn, max_steps)
        self.traffic_density = float(config.get("traffic_density", 0.5))
        self.env.configure_traffic(self.traffic_flow, self.traffic_density)
        mapping = DEFAULT_ROUTE_MAPPING_2LANES if self.num_lanes == 2 else DEFAULT_ROUTE_MAPPING_3LANES
        self.env.configure_routes([(s, e) for s, ends in mapping.items() for e in ends])
        reward_cfg = config.get("reward_config", None)
        if reward_cfg is None:
            reward_cfg = DEFAULT_REWARD_CONFIG.get("reward_config", {})
        if isinstance(reward_cfg, dict):
            _apply_reward_config(self.env, reward_cfg)
        self._cars = None            # env.py:152,155,184 snapshot env.cars / env.traffic_cars eagerly; here the device
        self._traffic_cars = None    # round trip (a sync + ~20 small copies) is paid only if somebody looks
        self.reset()

    @staticmethod
    def _default_routes(num_agents: int, num_lanes: int):
        mapping = DEFAULT_ROUTE_MAPPING_2LANES if num_lanes == 2 else DEFAULT_ROUTE_MAPPING_3LANES
        all_routes = [(s, e) for s, ends in mapping.items() for e in ends]
        return [all_routes[i % len(all_routes)] for i in range(num_agents)]

    def reset(self):
        self.env.reset()
        for i in range(self.num_agents):
            start_id, end_id = self.ego_routes[i]
            self.env.add_car_with_route(start_id, end_id)
        self._cars = self.env.cars                      # reset-time snapshot, as env.py:152 (resets are rare)
        self._traffic_cars = None
        obs = self._collect_obs()
        if self.traffic_flow:
            return obs[0], {}
        return obs, {}

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

    def _collect_obs(self) -> np.ndarray:
        return np.asarray(self.env.get_observations(), dtype=np.float32)

    def step(self, actions: Union[np.ndarray, List[List[float]], List[float]], dt: float = 1.0 / 60.0):
        actions = np.asarray(actions, dtype=np.float32)
        if self.traffic_flow:
            actions = actions.reshape(1, 2)
        elif actions.ndim == 1:
            if actions.size == 2 and self.num_agents == 1:
                actions = actions.reshape(1, 2)
            else:
                raise ValueError(f"Expected actions shape (N,2) for multi-agent, got {actions.shape}")
        res = self.env.step(actions[:, 0].tolist(), actions[:, 1].tolist(), float(dt))
        self._traffic_cars = None                       # refetched on access (state cannot change between steps)
        obs = np.asarray(res.obs, dtype=np.float32)
        rewards = np.asarray(res.rewards, dtype=np.float32)
        terminated, truncated = bool(res.terminated), bool(res.truncated)
        collisions = {int(res.agent_ids[i]): str(res.status[i]) for i in range(len(res.status))}
        info = {
            "step": int(res.step),
            "rewards": rewards.tolist() if not self.traffic_flow else float(rewards[0]) if len(rewards) else 0.0,
            "collisions": collisions,
            "agents_alive": int(res.agents_alive),
            "terminated": terminated,
            "truncated": truncated,
            "done": list(res.done),
            "status": list(res.status),
        }
        if self.traffic_flow:
            return obs[0], float(rewards[0]) if len(rewards) else 0.0, terminated, truncated, info
        return obs, rewards, terminated, truncated, info

    def render(self, show_lane_ids: bool | None = None, show_lidar: bool | None = None):
        return self.env.render()   # headless: an rgb array instead of the reference's Windows/GLFW window

    def close(self):
        if getattr(self.env, "_benv", None) is not None:
            self.env._benv.close()
            self.env._benv = None
