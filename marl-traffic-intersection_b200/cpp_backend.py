"""Backend loader with the surface of the reference's cpp_backend.py (/root/reference/cpp_backend.py:15-66):
``has_cpp_backend()`` and the ``IntersectionEnv`` factory.  Where the reference lazily imports the pybind11
module ``MARLEnv``, this loads libisx_b200.so and hands out a one-env view with MARLEnv.IntersectionEnv's
methods (bindings.cpp:58-83), so env.py-style code keeps working unchanged."""
from __future__ import annotations

import os
from typing import List, Tuple

import numpy as np

from . import _lib
from .utils import STATUS_NAMES, reward_vector


def has_cpp_backend() -> bool:
    return os.path.exists(_lib.LIB_PATH)


def _require():
    if not has_cpp_backend():
        raise RuntimeError("libisx_b200.so backend not available – build it first (python -m marl_traffic_intersection_b200.build).")
    return _lib.load_library()


class State:                     # bindings.cpp:14-19
    def __init__(self, x=0.0, y=0.0, v=0.0, heading=0.0):
        self.x, self.y, self.v, self.heading = float(x), float(y), float(v), float(heading)


class Car:                       # read-only snapshot with the fields bindings.cpp:21-31 exposes
    def __init__(self, cs=None, path=None):
        self.state = State()
        self.length, self.width = 54.0, 24.0
        self.alive, self.intention, self.path_index = True, 0, 0
        self.path: List[Tuple[float, float]] = path or []
        if cs is not None:
            self.state = State(cs.x, cs.y, cs.v, cs.heading)
            self.alive, self.intention, self.path_index = bool(cs.alive), int(cs.intention), int(cs.path_index)


class Lidar:                     # bindings.cpp:85-93 (default-constructed: 72 beams, Lidar.h:11-14)
    def __init__(self):
        self.rays, self.fov_deg, self.max_dist, self.step_size = 72, 360.0, 250.0, 4.0
        self.distances = [self.max_dist] * self.rays
        step = self.fov_deg / (self.rays - 1)
        self.rel_angles = [float(np.float32((np.float32(-180.0) + np.float32(i) * np.float32(step)) * np.float32(np.pi) / np.float32(180.0))) for i in range(self.rays)]

    def normalized(self):
        inv = (1.0 / self.max_dist) if self.max_dist > 0 else 0.0
        return [d * inv for d in self.distances]


class RewardConfig:              # Reward.h:5-14
    def __init__(self):
        (self.k_prog, self.v_min_ms, self.k_stuck, self.k_cv, self.k_co, self.k_succ, self.k_sm, self.alpha) = reward_vector(None)


class StepResult:                # Reward.h:16-29
    def __init__(self):
        self.obs, self.rewards, self.done, self.status, self.agent_ids = [], [], [], [], []
        self.agents_alive, self.terminated, self.truncated, self.step = 0, False, False, 0


class IntersectionEnv:
    """One env instance with MARLEnv.IntersectionEnv's call sequence: configure*, reset(), add_car_with_route()...,
    step().  The device handle is (re)built lazily at the first step / get_observations after the cars are added,
    because the batched SoA needs the agent count up front."""

    def __init__(self, num_lanes: int = 3):
        _require()
        self.num_lanes = int(num_lanes)
        self.reward_config = RewardConfig()
        self._use_team, self._respawn, self._max_steps = False, True, 2000
        self._traffic, self._density = False, 0.5
        self._traffic_routes = None
        self._pending: List[Tuple[str, str]] = []
        self._built_key = None
        self._benv = None
        self.seed = 0
        self.lidar_rays = 96
        self.step_count = 0

    def configure(self, use_team, respawn, max_steps):
        self._use_team, self._respawn, self._max_steps = bool(use_team), bool(respawn), int(max_steps)

    def configure_traffic(self, enabled, density):
        self._traffic, self._density = bool(enabled), max(0.0, float(density))

    def configure_routes(self, routes):
        self._traffic_routes = [(str(a), str(b)) for a, b in routes]

    def reset(self):
        self._pending = []
        self.step_count = 0
        self._needs_reset = True

    def add_car_with_route(self, start_id, end_id):
        lib = _lib.load_library()
        rc = lib.isx_route(self.num_lanes, str(start_id).encode(), str(end_id).encode(), None, None, None, None, None)
        if rc == _lib.E_ROUTE_START:
            return                                   # silent no-op, IntersectionEnv.cpp:79-82
        if rc == _lib.E_ROUTE_END:
            raise IndexError(f"unknown lane id {end_id!r}")   # std::out_of_range via .at(), RouteGen.cpp:120
        self._pending.append((str(start_id), str(end_id)))

    def _ensure(self):
        from .batched import BatchedIntersectionEnv
        rc = self.reward_config
        key = (tuple(self._pending), self._use_team, self._respawn, self._max_steps, self._traffic, self._density,
               tuple(self._traffic_routes or ()), (rc.k_prog, rc.v_min_ms, rc.k_stuck, rc.k_cv, rc.k_co, rc.k_succ, rc.k_sm, rc.alpha),
               self.seed, self.lidar_rays)
        if self._benv is None or key != self._built_key:
            if self._benv is not None:
                self._benv.close()
            if not self._pending:
                raise RuntimeError("no cars: call add_car_with_route() after reset()")
            self._benv = BatchedIntersectionEnv({
                "num_envs": 1, "num_agents": len(self._pending), "num_lanes": self.num_lanes, "ego_routes": self._pending,
                "use_team_reward": self._use_team, "respawn_enabled": self._respawn, "max_steps": self._max_steps,
                "traffic_flow": self._traffic, "traffic_density": self._density, "traffic_routes": self._traffic_routes,
                "reward_config": {"progress_scale": rc.k_prog, "stuck_speed_threshold": rc.v_min_ms, "stuck_penalty": rc.k_stuck,
                                  "crash_vehicle_penalty": rc.k_cv, "crash_object_penalty": rc.k_co, "success_reward": rc.k_succ,
                                  "action_smoothness_scale": rc.k_sm, "team_alpha": rc.alpha},
                "seed": self.seed, "lidar_rays": self.lidar_rays,
            })
            self._built_key = key
            self._needs_reset = False
        elif getattr(self, "_needs_reset", False):
            self._benv.reset()
            self._needs_reset = False
        return self._benv

    def get_observations(self):
        b = self._ensure()
        return b.buf["obs"][0].cpu().numpy().tolist()

    @property
    def cars(self):
        b = self._ensure()
        egos, _, _, _, _ = b.get_env_state(0)
        return [Car(egos[i]) for i in range(b.num_agents)]

    @property
    def traffic_cars(self):
        b = self._ensure()
        _, npcs, n, _, _ = b.get_env_state(0)
        return [Car(npcs[i]) for i in range(n)]

    def step(self, throttles, steerings, dt=1.0 / 60.0):
        b = self._ensure()
        n = b.num_agents
        a = np.zeros((1, n, 2), np.float32)         # missing actions default to 0, IntersectionEnv.cpp:153-154
        th = list(throttles)[:n]
        st = list(steerings)[:n]
        a[0, : len(th), 0] = th
        a[0, : len(st), 1] = st
        obs, rew, done, status, term, trunc = b.step_host(a, dt)
        r = StepResult()
        r.obs = obs[0].tolist()
        r.rewards = rew[0].tolist()
        r.done = [int(x) for x in done[0]]
        r.status = [STATUS_NAMES[int(x)] for x in status[0]]
        r.agent_ids = list(range(1, n + 1))          # IntersectionEnv.cpp:130: 1..N after every reset
        hv = b._host_views()                           # pinned views filled by the same step: no extra device round trips
        r.agents_alive = int(hv["agents_alive"][0])
        r.terminated, r.truncated = bool(term[0]), bool(trunc[0])
        r.step = int(hv["step"][0])
        self.step_count = r.step
        return r

    def render(self, show_lane_ids=False, show_lidar=False):
        """No window (the reference's is Windows/GLFW); returns the headless picture instead:
        uint8 ndarray [750, 750, 3] from isx_render, or None before the first reset."""
        if self._benv is None:
            return None
        return self._benv.render(0).cpu().numpy()

    def window_should_close(self):
        return True

    def poll_events(self):
        return None

    def key_pressed(self, glfw_key):
        return False
