"""Backend loader with the surface of the reference's cpp_backend.py (/root/reference/cpp_backend.py:15-66):
``has_cpp_backend()`` and the ``IntersectionEnv`` / ``Car`` / ``State`` / ``Lidar`` factories.  Where the reference lazily
imports the pybind11 module ``MARLEnv``, this loads libisx_b200.so and hands out a one-env view with
MARLEnv.IntersectionEnv's attributes and methods (bindings.cpp:58-83), so the reference's own env.py runs on top of it
unchanged (tests/test_gpu_dropin.py does exactly that: ``sys.modules["cpp_backend"] = this module``)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _lib
from .utils import STATUS_NAMES, reward_vector


def has_cpp_backend() -> bool:
    return os.path.exists(_lib.LIB_PATH)


def _require():
    if not has_cpp_backend():
        raise RuntimeError("libisx_b200.so backend not available – build it first (python -m marl_traffic_intersection_b200.build).")
    return _lib.load_library()


_CS_FLOATS = ("x", "y", "v", "heading", "acc", "steer", "prev_dist", "prev_a0", "prev_a1")
_path_cache: Dict[Tuple[int, str, str], List[Tuple[float, float]]] = {}


def _route_path(num_lanes: int, start: str, end: str) -> List[Tuple[float, float]]:
    """Car.path of a car on (start, end): the 160 way-points of RouteGen.cpp:111-205 as the library tabulates them."""
    key = (int(num_lanes), str(start), str(end))
    if key not in _path_cache:
        xy = (C.c_float * (2 * 160))()
        n = _require().isx_route(key[0], key[1].encode(), key[2].encode(), xy, None, None, None, None)
        _path_cache[key] = [(float(xy[2 * i]), float(xy[2 * i + 1])) for i in range(max(n, 0))]
    return list(_path_cache[key])


class State:                     # bindings.cpp:14-19
    def __init__(self, x=0.0, y=0.0, v=0.0, heading=0.0):
        self.x, self.y, self.v, self.heading = float(x), float(y), float(v), float(heading)


class Car:
    """By-value snapshot of one car with the fields bindings.cpp:21-31 exposes.  Like the pybind object it also carries the
    members the binding hides (acc, steering angle, reward bookkeeping, Car.h:23-37), so a Car taken from ``get_state()``
    and handed back to ``set_state()`` restores the car completely."""

    device = 0                   # GPU that evaluates update() / check_collision()

    def __init__(self, cs=None, path=None):
        self.state = State()
        self.length, self.width = 54.0, 24.0
        self.alive, self.intention, self.path_index = True, 0, 0
        self.path: List[Tuple[float, float]] = path or []
        self._acc = self._steer = self._prev_dist = self._prev_a0 = self._prev_a1 = 0.0
        self._route, self._uid = -1, 0
        if cs is not None:
            self.state = State(cs.x, cs.y, cs.v, cs.heading)
            self.alive, self.intention, self.path_index = bool(cs.alive), int(cs.intention), int(cs.path_index)
            self._acc, self._steer = float(cs.acc), float(cs.steer)
            self._prev_dist, self._prev_a0, self._prev_a1 = float(cs.prev_dist), float(cs.prev_a0), float(cs.prev_a1)
            self._route, self._uid = int(cs.route), int(cs.uid)

    def _record(self) -> "_lib.CarState":
        s = self.state
        return _lib.CarState(s.x, s.y, s.v, s.heading, self._acc, self._steer, self._prev_dist, self._prev_a0, self._prev_a1,
                             int(self.path_index), int(self._route), int(bool(self.alive)), int(self._uid), int(self.intention))

    def update(self, throttle, steer_input, dt):
        """Car::update (Car.cpp:9-40), evaluated on the GPU by the device function the step kernel uses."""
        lib, cs = _require(), self._record()
        _lib.check(lib, lib.isx_car_update(self.device, C.byref(cs), C.c_float(throttle), C.c_float(steer_input), C.c_float(dt)))
        self.state = State(cs.x, cs.y, cs.v, cs.heading)
        self._acc, self._steer = float(cs.acc), float(cs.steer)

    def check_collision(self, other: "Car") -> bool:
        """Car::check_collision (Car.cpp:105-141): SAT of the two 54x24 rectangles, touching counts."""
        lib, a, b, out = _require(), self._record(), other._record(), C.c_int32(0)
        _lib.check(lib, lib.isx_car_check_collision(self.device, C.byref(a), C.byref(b), C.byref(out)))
        return bool(out.value)


class Lidar:                     # bindings.cpp:85-93; default-constructed = 72 beams (Lidar.h:11-14)
    def __init__(self, rays: int = 72, distances=None):
        self.rays, self.fov_deg, self.max_dist, self.step_size = int(rays), 360.0, 250.0, 4.0
        self.distances = [self.max_dist] * self.rays if distances is None else [float(d) for d in distances]
        f = np.float32
        step = f(self.fov_deg) / f(self.rays - 1) if self.rays > 1 else f(0.0)          # Lidar.cpp:6-13, float32 throughout
        self.rel_angles = [float((f(-180.0) + f(i) * step) * f(np.pi) / f(180.0)) for i in range(self.rays)]

    def normalized(self):
        inv = np.float32(1.0) / np.float32(self.max_dist) if self.max_dist > 0 else np.float32(0.0)
        return [float(np.float32(d) * inv) for d in self.distances]


class RewardConfig:              # Reward.h:5-14
    def __init__(self):
        (self.k_prog, self.v_min_ms, self.k_stuck, self.k_cv, self.k_co, self.k_succ, self.k_sm, self.alpha) = reward_vector(None)

    def _vector(self):
        return (float(self.k_prog), float(self.v_min_ms), float(self.k_stuck), float(self.k_cv), float(self.k_co),
                float(self.k_succ), float(self.k_sm), float(self.alpha))


class StepResult:                # Reward.h:16-29
    def __init__(self):
        self.obs, self.rewards, self.done, self.status, self.agent_ids = [], [], [], [], []
        self.agents_alive, self.terminated, self.truncated, self.step = 0, False, False, 0


class EnvState:                  # EnvState.h:9-15, bindings.cpp:56-62
    def __init__(self):
        self.cars: List[Car] = []
        self.traffic_cars: List[Car] = []
        self.agent_ids: List[int] = []
        self.next_agent_id = 1
        self.step_count = 0


class IntersectionEnv:
    """One env instance with MARLEnv.IntersectionEnv's call sequence: configure*, reset(), add_car_with_route()...,
    step().  The device handle is built lazily at the first step / get_observations after the cars are added (the batched
    SoA needs the agent count up front) and rebuilt only when something that shapes the buffers changes (cars, traffic
    on/off, traffic routes); reward weights, configure() and the traffic density are applied to the live handle."""

    def __init__(self, num_lanes: int = 3):
        _require()
        self.num_lanes = int(num_lanes)
        self.reward_config = RewardConfig()
        self._use_team, self._respawn, self._max_steps = False, True, 2000
        self._traffic, self._density = False, 0.5
        self._traffic_routes = None
        self._pending: List[Tuple[str, str]] = []
        self._hard_key = self._soft_key = None
        self._benv = None
        self._needs_reset = False
        self.seed = 0
        self.lidar_rays = 96         # add_car_with_route gives every ego a 96-beam lidar (IntersectionEnv.cpp:112-128)
        self._step_count = 0

    # ------------------------------------------------------------------ configuration (bindings.cpp:64-69)
    def configure(self, use_team, respawn, max_steps):
        self._use_team, self._respawn, self._max_steps = bool(use_team), bool(respawn), int(max_steps)

    def configure_traffic(self, enabled, density):
        self._traffic, self._density = bool(enabled), max(0.0, float(density))

    def configure_routes(self, routes):
        self._traffic_routes = [(str(a), str(b)) for a, b in routes]

    def reset(self):
        self._pending = []
        self._step_count = 0
        self._needs_reset = True

    def add_car_with_route(self, start_id, end_id):
        lib = _lib.load_library()
        rc = lib.isx_route(self.num_lanes, str(start_id).encode(), str(end_id).encode(), None, None, None, None, None)
        if rc == _lib.E_ROUTE_START:
            return                                   # silent no-op, IntersectionEnv.cpp:79-82
        if rc == _lib.E_ROUTE_END:
            raise IndexError(f"unknown lane id {end_id!r}")   # std::out_of_range via .at(), RouteGen.cpp:120
        self._pending.append((str(start_id), str(end_id)))

    def _soft(self):
        return (self._use_team, self._respawn, self._max_steps, self._density, self.reward_config._vector())

    def _ensure(self):
        from .batched import BatchedIntersectionEnv
        hard = (tuple(self._pending), self._traffic, tuple(self._traffic_routes or ()), self.seed, self.num_lanes)
        if self._benv is None or hard != self._hard_key:
            if self._benv is not None:
                self._benv.close()
            if not self._pending:
                raise RuntimeError("no cars: call add_car_with_route() after reset()")
            rc = self.reward_config
            self._benv = BatchedIntersectionEnv({
                "num_envs": 1, "num_agents": len(self._pending), "num_lanes": self.num_lanes, "ego_routes": self._pending,
                "use_team_reward": self._use_team, "respawn_enabled": self._respawn, "max_steps": self._max_steps,
                "traffic_flow": self._traffic, "traffic_density": self._density, "traffic_routes": self._traffic_routes,
                "reward_config": rc._vector(), "seed": self.seed, "lidar_rays": self.lidar_rays,
            })
            self._hard_key, self._soft_key = hard, self._soft()
            self._needs_reset = False
            return self._benv
        b = self._benv
        soft = self._soft()
        if soft != self._soft_key:                   # live settings: no rebuild
            b.configure(self._use_team, self._respawn, self._max_steps)
            b.set_traffic_density(self._density)
            b.set_reward_config(self.reward_config._vector())
            self._soft_key = soft
        if self._needs_reset:
            if b.lidar_rays != self.lidar_rays:      # reset() + add_car_with_route rebuild the 96-beam lidars (:112-128)
                b.set_lidar_rays(self.lidar_rays)
            b.reset()
            self._needs_reset = False
        return b

    # ------------------------------------------------------------------ state views (def_readwrite copies, bindings.cpp:60-63)
    def get_observations(self):
        b = self._ensure()
        return b.buf["obs"][0].cpu().numpy().tolist()

    def _ego_path(self, i):
        return _route_path(self.num_lanes, *self._benv.ego_routes[i])

    def _npc_path(self, route):
        tr = self._benv.traffic_routes
        return _route_path(self.num_lanes, *tr[route]) if 0 <= route < len(tr) else []

    @property
    def cars(self):
        b = self._ensure()
        egos, _, _, _, _ = b.get_env_state(0)
        return [Car(egos[i], self._ego_path(i)) for i in range(b.num_agents)]

    @property
    def traffic_cars(self):
        b = self._ensure()
        _, npcs, n, _, _ = b.get_env_state(0)
        return [Car(npcs[i], self._npc_path(int(npcs[i].route))) for i in range(n)]

    @property
    def lidars(self):
        """One Lidar per ego with the distances of the last step (bindings.cpp:62; 250 = nothing within range)."""
        b = self._ensure()
        hits = b.buf["lidar_hit"][0].cpu().numpy()
        R = b.lidar_rays
        return [Lidar(R, [4.0 * k if k else 250.0 for k in hits[a, :R].tolist()]) for a in range(b.num_agents)]

    @property
    def step_count(self):
        return self._step_count

    @step_count.setter
    def step_count(self, v):
        self._step_count = int(v)
        if self._benv is not None and not self._needs_reset:
            egos, npcs, n, _, tick = self._benv.get_env_state(0)
            self._benv.set_env_state(0, egos, npcs, n, int(v), tick)

    def get_state(self) -> EnvState:
        """IntersectionEnv::get_state (IntersectionEnv.cpp:394-402): a by-value snapshot for MCTS-style rollbacks."""
        b = self._ensure()
        egos, npcs, n, sc, _ = b.get_env_state(0)
        s = EnvState()
        s.cars = [Car(egos[i], self._ego_path(i)) for i in range(b.num_agents)]
        s.traffic_cars = [Car(npcs[i], self._npc_path(int(npcs[i].route))) for i in range(n)]
        s.agent_ids = list(range(1, b.num_agents + 1))           # :130
        s.next_agent_id = b.num_agents + 1
        s.step_count = int(sc)
        return s

    def set_state(self, state: EnvState):
        """IntersectionEnv::set_state (IntersectionEnv.cpp:404-416), including what it does to the lidars: they are
        replaced by default-constructed ones — 72 beams, every distance 250 — so from here until the next reset() the env
        observes with 72 beams (obs[31:103]; obs[103:] stays 0).  The number of ego cars must match the live env."""
        b = self._ensure()
        if len(state.cars) != b.num_agents:
            raise ValueError(f"set_state: {len(state.cars)} ego cars for an env with {b.num_agents} (rebuild with reset() + add_car_with_route())")
        if len(state.traffic_cars) > b.npc_capacity or (state.traffic_cars and not b.traffic_flow):
            raise ValueError("set_state: more traffic cars than NPC slots")
        egos = (_lib.CarState * b.num_agents)(*[c._record() for c in state.cars])
        for i in range(b.num_agents):
            egos[i].route = i
        nn = len(state.traffic_cars)
        npcs = (_lib.CarState * max(nn, 1))(*[c._record() for c in state.traffic_cars])
        _, _, _, _, tick = b.get_env_state(0)                      # the RNG position is not part of EnvState (EnvState.h:9-15)
        b.set_env_state(0, egos, npcs, nn, int(state.step_count), tick)
        b.set_lidar_rays(72)                                       # lidars.clear(); lidars.resize(n) -> Lidar() (:411-415)
        self._step_count = int(state.step_count)

    # ------------------------------------------------------------------ step (bindings.cpp:76)
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
        self._step_count = r.step
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
