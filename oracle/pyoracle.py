"""oracle/pyoracle.py — TEST INFRASTRUCTURE (not product code).

ctypes front-end shared by the two CPU checkers:

* ``RefEnv``    -> oracle/_ref/libisx_ref.so  (the reference's own C++, ref_driver.cpp)
* ``OracleEnv`` -> oracle/libisx_oracle.so    (the plain-C restatement, isx_oracle.c)

Both expose the same methods so a parity test can be parametrised over them.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libisx_ref.so")
ORACLE_SO = os.path.join(HERE, "libisx_oracle.so")

OBS_DIM = 127
PATH_LEN = 160
STATUS_NAMES = ["ALIVE", "DEAD", "SUCCESS", "CRASH_WALL", "CRASH_LINE", "CRASH_CAR"]

CAR_DTYPE = np.dtype(
    [
        ("x", "<f4"), ("y", "<f4"), ("v", "<f4"), ("heading", "<f4"),
        ("acc", "<f4"), ("steer", "<f4"),
        ("prev_dist", "<f4"), ("prev_a0", "<f4"), ("prev_a1", "<f4"),
        ("path_index", "<i4"), ("route", "<i4"), ("alive", "<i4"),
        ("uid", "<u4"), ("intention", "<i4"),
    ],
    align=False,
)
assert CAR_DTYPE.itemsize == 56

EVENTS_DTYPE = np.dtype(
    [
        ("rng_draws", "<i4"), ("spawn_route", "<i4"), ("spawned", "<i4"),
        ("removed_mask", "<u4"), ("collided_mask", "<u4"), ("npc_count", "<i4"),
    ]
)
assert EVENTS_DTYPE.itemsize == 24

DEFAULT_REWARD = (10.0, 1.0, -0.01, -10.0, -5.0, 10.0, -0.02, 0.2)

# /root/reference/utils.py:29-52 (dict iteration order)
ROUTES_3LANES = [
    ("IN_1", "OUT_4"), ("IN_2", "OUT_8"), ("IN_3", "OUT_12"), ("IN_4", "OUT_7"),
    ("IN_5", "OUT_11"), ("IN_6", "OUT_3"), ("IN_7", "OUT_10"), ("IN_8", "OUT_2"),
    ("IN_9", "OUT_6"), ("IN_10", "OUT_1"), ("IN_11", "OUT_5"), ("IN_12", "OUT_9"),
]
ROUTES_2LANES = [
    ("IN_1", "OUT_3"), ("IN_2", "OUT_6"), ("IN_3", "OUT_5"), ("IN_4", "OUT_8"),
    ("IN_6", "OUT_2"), ("IN_7", "OUT_1"), ("IN_8", "OUT_4"),
]


def default_routes(num_lanes: int) -> List[Tuple[str, str]]:
    return list(ROUTES_2LANES if num_lanes == 2 else ROUTES_3LANES)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _strarr(items: Sequence[str]):
    arr = (C.c_char_p * len(items))()
    arr[:] = [s.encode() for s in items]
    return arr


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def have_oracle() -> bool:
    return os.path.exists(ORACLE_SO)


_libs = {}


def _load(path: str, prefix: str):
    key = (path, prefix)
    if key in _libs:
        return _libs[key]
    lib = C.CDLL(path)

    class _Missing:  # a library may export only a subset (the product's host-unit hooks do)
        restype = None
        argtypes = None

    def f(n):
        try:
            return getattr(lib, prefix + n)
        except AttributeError:
            return _Missing()
    f("create").restype = C.c_void_p
    f("create").argtypes = [C.c_int]
    f("destroy").argtypes = [C.c_void_p]
    f("configure").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    f("configure_traffic").argtypes = [C.c_void_p, C.c_int, C.c_float]
    f("configure_routes").argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]
    f("set_reward").argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    f("set_ego_routes").argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]
    f("set_lidar_rays").argtypes = [C.c_void_p, C.c_int]
    f("reset").argtypes = [C.c_void_p]
    f("reset").restype = C.c_int
    f("seed").argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32]
    f("tick").argtypes = [C.c_void_p]
    f("tick").restype = C.c_uint32
    f("num_agents").argtypes = [C.c_void_p]
    f("num_npcs").argtypes = [C.c_void_p]
    f("step_count").argtypes = [C.c_void_p]
    f("set_step_count").argtypes = [C.c_void_p, C.c_int]
    f("get_obs").argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    f("step").argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_float,
                          C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                          C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    f("step").restype = C.c_int
    f("get_events").argtypes = [C.c_void_p, C.c_void_p]
    f("get_egos").argtypes = [C.c_void_p, C.c_void_p]
    f("get_npcs").argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    f("get_lidar").argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_int]
    f("set_egos").argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    f("set_npcs").argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    f("rollout").argtypes = [C.c_void_p, C.c_int, C.c_float, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    f("rollout").restype = C.c_longlong
    f("route").argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.POINTER(C.c_float), C.POINTER(C.c_int),
                           C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    f("on_road").argtypes = [C.c_int, C.c_float, C.c_float]
    f("yellow").argtypes = [C.c_int, C.c_float, C.c_float]
    f("is_line").argtypes = [C.c_int, C.c_int, C.c_int]
    f("road_map").argtypes = [C.c_int, C.c_void_p]
    f("line_map").argtypes = [C.c_int, C.c_void_p]
    f("car_update").argtypes = [C.POINTER(C.c_float), C.c_float, C.c_float, C.c_float]
    f("collide").argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float)]
    f("corners").argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float)]
    f("lidar").argtypes = [C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float)]
    for n in ("libm_sincosf",):
        f(n).argtypes = [C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    f("libm_tanf").argtypes = [C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float)]
    for n in ("libm_atan2f", "libm_hypotf", "libm_fmodf"):
        f(n).argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float)]
    _libs[key] = lib
    return lib


class _Unit:
    """Stateless probes of single functions (route generation, geometry, car, lidar, libm)."""

    def __init__(self, lib, prefix):
        self._lib = lib
        self._p = prefix

    def _f(self, n):
        return getattr(self._lib, self._p + n)

    def route(self, num_lanes: int, start: str, end: str):
        path = np.zeros((PATH_LEN, 2), np.float32)
        intent = C.c_int(0)
        sx, sy, sh = C.c_float(), C.c_float(), C.c_float()
        n = self._f("route")(num_lanes, start.encode(), end.encode(), _fp(path), C.byref(intent),
                             C.byref(sx), C.byref(sy), C.byref(sh))
        if n < 0:
            return n, None, None, None
        return n, path[:n].copy(), intent.value, np.array([sx.value, sy.value, sh.value], np.float32)

    def on_road(self, lanes, x, y):
        return bool(self._f("on_road")(lanes, float(x), float(y)))

    def yellow(self, lanes, x, y):
        return bool(self._f("yellow")(lanes, float(x), float(y)))

    def is_line(self, lanes, x, y):
        return bool(self._f("is_line")(lanes, int(x), int(y)))

    def road_map(self, lanes):
        out = np.zeros((750, 750), np.uint8)
        self._f("road_map")(lanes, out.ctypes.data)
        return out

    def line_map(self, lanes):
        out = np.zeros((750, 750), np.uint8)
        self._f("line_map")(lanes, out.ctypes.data)
        return out

    def car_update(self, s6, thr, st, dt):
        s = np.array(s6, np.float32)
        self._f("car_update")(_fp(s), float(thr), float(st), float(dt))
        return s

    def collide(self, a3, b3):
        a = np.array(a3, np.float32)
        b = np.array(b3, np.float32)
        return bool(self._f("collide")(_fp(a), _fp(b)))

    def corners(self, a3):
        a = np.array(a3, np.float32)
        o = np.zeros(8, np.float32)
        self._f("corners")(_fp(a), _fp(o))
        return o.reshape(4, 2)

    def std_sort(self, keys):
        """Rank order the library's neighbour sort leaves `keys` in: (perm[int32], heap_fallbacks or None)."""
        k = np.ascontiguousarray(keys, np.float32)
        perm = np.zeros(max(k.shape[0], 1), np.int32)
        f = self._f("std_sort")
        f.restype = C.c_int if self._p != "isxref_" else None
        r = f(_fp(k), k.shape[0], _ip(perm))
        return perm[: k.shape[0]].copy(), r

    def sort_adversary(self, n):
        out = np.zeros(n, np.float32)
        f = self._f("sort_adversary")
        f.restype = None
        f(int(n), _fp(out))
        return out

    def lidar(self, lanes, rays, self_pose, others):
        sp = np.array(self_pose, np.float32)
        ot = np.ascontiguousarray(np.array(others, np.float32).reshape(-1, 3))
        d = np.zeros(rays, np.float32)
        self._f("lidar")(lanes, rays, _fp(sp), _fp(ot), ot.shape[0], _fp(d))
        return d

    def sincosf(self, x):
        x = np.ascontiguousarray(x, np.float32)
        s = np.empty_like(x)
        c = np.empty_like(x)
        self._f("libm_sincosf")(_fp(x), x.size, _fp(s), _fp(c))
        return s, c

    def tanf(self, x):
        x = np.ascontiguousarray(x, np.float32)
        o = np.empty_like(x)
        self._f("libm_tanf")(_fp(x), x.size, _fp(o))
        return o

    def _bin(self, name, y, x):
        y = np.ascontiguousarray(y, np.float32)
        x = np.ascontiguousarray(x, np.float32)
        o = np.empty_like(x)
        self._f(name)(_fp(y), _fp(x), x.size, _fp(o))
        return o

    def atan2f(self, y, x):
        return self._bin("libm_atan2f", y, x)

    def hypotf(self, y, x):
        return self._bin("libm_hypotf", y, x)

    def fmodf(self, y, x):
        return self._bin("libm_fmodf", y, x)


class _EnvBase:
    """One CPU env instance with the surface env.py drives (env.py:111-131,147-208)."""

    _so = ""
    _prefix = ""

    def __init__(self, num_lanes: int = 3, ego_routes=None, use_team=False, respawn=True, max_steps=2000,
                 traffic=False, density=0.5, traffic_routes="default", reward=DEFAULT_REWARD, lidar_rays=96,
                 seed=0, env_id=0):
        self._lib = _load(self._so, self._prefix)
        self._h = C.c_void_p(self._f("create")(num_lanes))
        self.num_lanes = num_lanes
        self._f("configure")(self._h, int(use_team), int(respawn), int(max_steps))
        self._f("configure_traffic")(self._h, int(traffic), float(density))
        if traffic_routes == "default":
            traffic_routes = default_routes(num_lanes)  # env.py:118-123
        if traffic_routes is not None:
            self._f("configure_routes")(self._h, len(traffic_routes), _strarr([a for a, _ in traffic_routes]),
                                        _strarr([b for _, b in traffic_routes]))
        self.traffic_routes = traffic_routes
        r = np.array(reward, np.float32)
        self._f("set_reward")(self._h, _fp(r))
        if ego_routes is None:
            ego_routes = [("IN_6", "OUT_2")]
        self.ego_routes = list(ego_routes)
        self._f("set_ego_routes")(self._h, len(ego_routes), _strarr([a for a, _ in ego_routes]),
                                  _strarr([b for _, b in ego_routes]))
        self.lidar_rays = lidar_rays
        self._f("set_lidar_rays")(self._h, int(lidar_rays))
        self._f("seed")(self._h, int(seed), int(env_id), 0)
        self.n = self.reset()

    def _f(self, n):
        return getattr(self._lib, self._prefix + n)

    def close(self):
        if self._h:
            self._f("destroy")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self) -> int:
        n = self._f("reset")(self._h)
        if n < 0:
            raise IndexError("unknown end lane id")
        self.n = n
        return n

    def seed(self, seed, env_id=0, tick=0):
        self._f("seed")(self._h, int(seed), int(env_id), int(tick))

    @property
    def tick(self):
        return int(self._f("tick")(self._h))

    @property
    def step_count(self):
        return int(self._f("step_count")(self._h))

    @step_count.setter
    def step_count(self, v):
        self._f("set_step_count")(self._h, int(v))

    def obs(self):
        o = np.zeros((self.n, OBS_DIM), np.float32)
        self._f("get_obs")(self._h, _fp(o))
        return o

    def step(self, actions, dt=1.0 / 60.0):
        a = np.ascontiguousarray(np.asarray(actions, np.float32).reshape(-1, 2))
        th = np.ascontiguousarray(a[:, 0])
        st = np.ascontiguousarray(a[:, 1])
        n = self.n
        obs = np.zeros((n, OBS_DIM), np.float32)
        rew = np.zeros(n, np.float32)
        done = np.zeros(n, np.int32)
        status = np.zeros(n, np.int32)
        term, trunc, alive = C.c_int32(), C.c_int32(), C.c_int32()
        step = self._f("step")(self._h, _fp(th), _fp(st), a.shape[0], np.float32(dt), _fp(obs), _fp(rew), _ip(done),
                               _ip(status), C.byref(term), C.byref(trunc), C.byref(alive))
        return dict(obs=obs, reward=rew, done=done, status=status, terminated=bool(term.value),
                    truncated=bool(trunc.value), agents_alive=alive.value, step=step)

    def events(self):
        ev = np.zeros(1, EVENTS_DTYPE)
        self._f("get_events")(self._h, ev.ctypes.data)
        return ev[0]

    def egos(self):
        out = np.zeros(max(self.n, 1), CAR_DTYPE)
        n = self._f("get_egos")(self._h, out.ctypes.data)
        return out[:n].copy()

    def npcs(self, cap=64):
        out = np.zeros(cap, CAR_DTYPE)
        n = self._f("get_npcs")(self._h, out.ctypes.data, cap)
        return out[: min(n, cap)].copy()

    def lidar(self, agent):
        d = np.zeros(128, np.float32)
        n = self._f("get_lidar")(self._h, agent, _fp(d), 128)
        return d[:n].copy()

    def set_egos(self, cars):
        cars = np.ascontiguousarray(cars, CAR_DTYPE)
        self._f("set_egos")(self._h, cars.ctypes.data, cars.shape[0])

    def set_npcs(self, cars):
        cars = np.ascontiguousarray(cars, CAR_DTYPE)
        self._f("set_npcs")(self._h, cars.ctypes.data, cars.shape[0])

    def rollout(self, steps, dt=1.0 / 60.0):
        hist = np.zeros(6, np.int32)
        rs = C.c_double(0.0)
        n = self._f("rollout")(self._h, int(steps), np.float32(dt), _ip(hist), C.byref(rs))
        return int(n), hist, rs.value


class RefEnv(_EnvBase):
    _so = REF_SO
    _prefix = "isxref_"


class OracleEnv(_EnvBase):
    _so = ORACLE_SO
    _prefix = "isxo_"


def step_batch(envs, actions, dt=1.0 / 60.0, lidar=False, npc_cap=0, threads=0):
    """Steps a list of RefEnv objects with ONE library call on `threads` host threads (isxref_step_batch) and returns dense
    arrays over the env index: obs [E,N,127], reward, done, status [E,N], terminated, truncated, agents_alive, step [E],
    events [E] (EVENTS_DTYPE), optionally lidar_k [E,N,96] (hit sample index, 0 = none) and npc_pose [E,npc_cap,4]."""
    lib = _load(REF_SO, "isxref_")
    f = lib.isxref_step_batch
    f.restype = None
    f.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_float), C.c_int, C.c_float] + [C.c_void_p] * 11 + [C.c_int, C.c_int]
    E, N = len(envs), envs[0].n
    a = np.ascontiguousarray(actions, np.float32).reshape(E, N, 2)
    hs = (C.c_void_p * E)(*[e._h for e in envs])
    out = dict(obs=np.zeros((E, N, OBS_DIM), np.float32), reward=np.zeros((E, N), np.float32), done=np.zeros((E, N), np.int32),
               status=np.zeros((E, N), np.int32), terminated=np.zeros(E, np.int32), truncated=np.zeros(E, np.int32),
               agents_alive=np.zeros(E, np.int32), step=np.zeros(E, np.int32), events=np.zeros(E, EVENTS_DTYPE))
    lk = np.zeros((E, N, 96), np.uint8) if lidar else None
    pose = np.zeros((E, max(npc_cap, 1), 4), np.float32) if npc_cap > 0 else None
    f(hs, E, _fp(a), N, np.float32(dt), out["obs"].ctypes.data, out["reward"].ctypes.data, out["done"].ctypes.data,
      out["status"].ctypes.data, out["terminated"].ctypes.data, out["truncated"].ctypes.data, out["agents_alive"].ctypes.data,
      out["step"].ctypes.data, out["events"].ctypes.data, lk.ctypes.data if lidar else None, pose.ctypes.data if pose is not None else None,
      int(npc_cap), int(threads or os.cpu_count() or 1))
    if lidar:
        out["lidar_k"] = lk
    if pose is not None:
        out["npc_pose"] = pose
    return out


def philox_actions_batch(seeds, env_ids, ticks, n_agents: int) -> np.ndarray:
    """philox_actions for many envs at once: seeds / env_ids / ticks are length-E sequences -> [E, n_agents, 2]."""
    E = len(env_ids)
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    env = np.repeat(np.asarray(env_ids, np.uint32), n_agents)
    tick = np.repeat(np.asarray(ticks, np.uint32), n_agents)
    seeds = np.asarray(seeds, np.uint64)
    k0 = np.repeat((seeds & np.uint64(0xFFFFFFFF)).astype(np.uint32), n_agents)
    k1 = np.repeat((seeds >> np.uint64(32)).astype(np.uint32), n_agents)
    c = [env, tick, np.tile(np.arange(n_agents, dtype=np.uint32), E), np.full(E * n_agents, 0x41435431, np.uint32)]
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c[1] ^ k0
            n1 = p1.astype(np.uint32)
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c[3] ^ k1
            n3 = p0.astype(np.uint32)
            c = [n0, n1, n2, n3]
            k0 = k0 + np.uint32(0x9E3779B9)
            k1 = k1 + np.uint32(0xBB67AE85)
    thr = (c[0] >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 8388608.0) - np.float32(1.0)
    st = (c[1] >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 8388608.0) - np.float32(1.0)
    return np.stack([thr, st], axis=1).reshape(E, n_agents, 2)


def unit_of(path: str, prefix: str) -> _Unit:
    """Unit-probe view of any library exporting the <prefix>route / on_road / lidar ... subset."""
    return _Unit(_load(path, prefix), prefix)


def ref_unit() -> _Unit:
    return _Unit(_load(REF_SO, "isxref_"), "isxref_")


def oracle_unit() -> _Unit:
    return _Unit(_load(ORACLE_SO, "isxo_"), "isxo_")


def philox_actions(seed: int, env_id: int, tick: int, n_agents: int) -> np.ndarray:
    """Host copy of the action stream defined in oracle/philox.h (numpy, vectorised over agents)."""
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    c = [np.full(n_agents, env_id, np.uint32), np.full(n_agents, tick, np.uint32),
         np.arange(n_agents, dtype=np.uint32), np.full(n_agents, 0x41435431, np.uint32)]
    k0 = np.uint32(seed & 0xFFFFFFFF)
    k1 = np.uint32((seed >> 32) & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c[1] ^ k0
            n1 = p1.astype(np.uint32)
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c[3] ^ k1
            n3 = p0.astype(np.uint32)
            c = [n0, n1, n2, n3]
            k0 = np.uint32((int(k0) + 0x9E3779B9) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + 0xBB67AE85) & 0xFFFFFFFF)
    thr = (c[0] >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 8388608.0) - np.float32(1.0)
    st = (c[1] >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 8388608.0) - np.float32(1.0)
    return np.stack([thr, st], axis=1)
