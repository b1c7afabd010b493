"""Batched front-end: E env instances stepping in lockstep on one B200 through libisx_b200.so.

Mirrors the reference's ``IntersectionEnv`` reset()/step() contract (/root/reference/env.py:147-208) with a
leading env dimension: ``reset() -> obs[E,N,127]``; ``step(actions[E,N,2], dt) -> obs, reward[E,N],
terminated[E], truncated[E], info``.  Every returned tensor is a NON-OWNING view of a device buffer that the
library overwrites on the next step (the reference returns by-value copies, bindings.cpp:60-62)."""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .utils import STATUS_NAMES, all_default_routes, default_ego_routes, reward_vector

_NP = {"f4": np.float32, "u1": np.uint8, "i4": np.int32, "u4": np.uint32}
_TORCH_VIEW = {"u4": "i4"}  # torch has no uint32 arithmetic; expose as int32 bits


class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch can wrap a raw device pointer without copying."""

    def __init__(self, ptr: int, shape: Tuple[int, ...], typestr: str, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<" + typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}
        self._owner = owner


def _strarr(items: Sequence[str]):
    arr = (C.c_char_p * max(len(items), 1))()
    for i, s in enumerate(items):
        arr[i] = s.encode()
    return arr


def reduce_stat_tensors(counters: torch.Tensor, reward_sum: torch.Tensor, group=None) -> Dict[str, Any]:
    """all-reduce(sum) of the int64[15] counter vector and — separately, as a double — of reward_sum[1] over the ranks of
    `group` (torch.distributed: NCCL for device tensors, gloo for host tensors), then the named totals.  Reduces IN PLACE.
    Without an initialised process group the local values are returned."""
    import torch.distributed as dist
    if counters.dtype != torch.int64 or reward_sum.dtype != torch.float64:
        raise TypeError("counters must be int64 and reward_sum float64 (a double's bit pattern must not ride in an integer sum)")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(reward_sum, op=dist.ReduceOp.SUM, group=group)
    c = counters.cpu().tolist()
    out = {k: c[i] for k, i in _lib.STAT_INDEX.items()}
    out["status_hist"] = {STATUS_NAMES[i]: c[i] for i in range(6)}
    out["reward_sum"] = float(reward_sum.cpu()[0])
    return out


class BatchedIntersectionEnv:
    """config keys (all optional): num_envs, num_agents, num_lanes, ego_routes, use_team_reward,
    respawn_enabled, max_steps, traffic_flow, traffic_density, traffic_routes, reward_config, lidar_rays (96|72),
    npc_capacity, seed, device, env_id_base, auto_reset.

    ``auto_reset`` (the reference has none; a caller of env.py resets after terminated|truncated): 0/False = call
    ``reset(mask)`` yourself; 1/True = the env is reset at the START of the step after a terminated|truncated step and that
    step's action already drives the new episode (the policy chose it from the terminal observation; the reset observation
    is never returned) — what the random-action benchmark uses; 2 = next-step reset as in Gymnasium's vector envs: the call
    after a terminated|truncated step only resets (action ignored, reward 0, flags 0) and returns the reset observation.

    ``config`` may also be a LIST of such dicts: a heterogeneous batch whose env range is the concatenation of the
    groups (isx_create_groups) — routes, lane count, traffic, reward weights and episode settings per group, one set of
    [sum(num_envs), N, ...] tensors.  ``group_ranges`` lists (first_env, num_envs) per group."""

    def __init__(self, config=None):
        cfgs = [dict(c) for c in config] if isinstance(config, (list, tuple)) else [dict(config or {})]
        if not cfgs:
            raise ValueError("empty config list")
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedIntersectionEnv needs a CUDA device: there is no CPU fallback")
        self._lib = _lib.load_library()
        dev = cfgs[0].get("device", None)
        self.device_index = torch.cuda.current_device() if dev is None else torch.device(dev).index or 0
        self.device = torch.device("cuda", self.device_index)
        arr = (_lib.Config * len(cfgs))()
        self._keep = []
        self.group_configs = []
        for g, cfg in enumerate(cfgs):
            self.group_configs.append(self._fill_config(arr[g], cfg))
        g0 = self.group_configs[0]
        self.num_envs = sum(gc["num_envs"] for gc in self.group_configs)
        self.num_agents = g0["num_agents"]
        self.num_lanes = g0["num_lanes"]
        self.lidar_rays = g0["lidar_rays"]
        self.max_steps = g0["max_steps"]
        self.ego_routes = g0["ego_routes"]
        self.traffic_routes = g0["traffic_routes"]
        self.traffic_flow = any(gc["traffic_flow"] for gc in self.group_configs)
        self._h = C.c_void_p()
        if len(cfgs) == 1:
            rc = self._lib.isx_create(C.byref(arr[0]), C.byref(self._h))
        else:
            rc = self._lib.isx_create_groups(arr, len(cfgs), C.byref(self._h))
        if rc == _lib.E_ROUTE_END:
            raise IndexError(self._lib.isx_last_error().decode())       # std::out_of_range -> IndexError in the reference
        _lib.check(self._lib, rc)
        caps = [gc["npc_capacity"] for gc in self.group_configs if gc["traffic_flow"]]
        self.npc_capacity = caps[0] if caps else 1
        self.group_ranges = []
        for g in range(len(cfgs)):
            a, b = C.c_int32(), C.c_int32()
            _lib.check(self._lib, self._lib.isx_group_range(self._h, g, C.byref(a), C.byref(b)))
            self.group_ranges.append((a.value, b.value))
        self._wrap_buffers()

    def _fill_config(self, c, cfg: Dict[str, Any]) -> Dict[str, Any]:
        """One isx_config from one config dict (defaults of env.py:41-78,111-145); returns the resolved settings."""
        num_lanes = int(cfg.get("num_lanes", 3))
        num_agents = int(cfg.get("num_agents", 1))
        routes = cfg.get("ego_routes", None)
        if routes is None:
            routes = default_ego_routes(num_agents, num_lanes)             # env.py:138-145
        routes = [(str(a), str(b)) for a, b in routes]
        if len(routes) != num_agents:
            raise ValueError(f"ego_routes has {len(routes)} entries for num_agents={num_agents}")
        troutes = cfg.get("traffic_routes", None)
        if troutes is None:
            troutes = all_default_routes(num_lanes)                        # env.py:118-123
        troutes = [(str(a), str(b)) for a, b in troutes]
        c.abi_version = _lib.ISX_ABI_VERSION
        c.device = self.device_index
        c.num_envs = int(cfg.get("num_envs", 1))
        c.num_agents = num_agents
        c.num_lanes = num_lanes
        c.lidar_rays = int(cfg.get("lidar_rays", 96))
        c.npc_capacity = int(cfg.get("npc_capacity", 16))
        c.use_team_reward = int(bool(cfg.get("use_team_reward", False)))
        c.respawn_enabled = int(bool(cfg.get("respawn_enabled", True)))
        c.max_steps = int(cfg.get("max_steps", 2000))
        c.traffic_flow = int(bool(cfg.get("traffic_flow", False)))
        c.traffic_density = float(cfg.get("traffic_density", 0.5))
        for i, v in enumerate(reward_vector(cfg.get("reward_config", None))):
            c.reward[i] = v
        keep = [_strarr([a for a, _ in routes]), _strarr([b for _, b in routes]),
                _strarr([a for a, _ in troutes]), _strarr([b for _, b in troutes])]
        self._keep.append(keep)
        c.ego_start, c.ego_end = keep[0], keep[1]
        c.num_traffic_routes = len(troutes)
        c.traffic_start, c.traffic_end = keep[2], keep[3]
        c.seed = int(cfg.get("seed", 0)) & 0xFFFFFFFFFFFFFFFF
        c.env_id_base = int(cfg.get("env_id_base", 0))
        c.auto_reset = int(cfg.get("auto_reset", 0))          # 0 off, 1/True reset-and-act, 2 next-step reset (isx.h)
        return dict(num_envs=c.num_envs, num_agents=num_agents, num_lanes=num_lanes, lidar_rays=c.lidar_rays,
                    max_steps=c.max_steps, ego_routes=routes, traffic_routes=troutes, traffic_flow=bool(c.traffic_flow),
                    npc_capacity=c.npc_capacity if c.npc_capacity > 0 else 16)

    # ------------------------------------------------------------------ buffers
    def _wrap_buffers(self):
        b = _lib.Buffers()
        _lib.check(self._lib, self._lib.isx_get_buffers(self._h, C.byref(b)))
        E, N, M = self.num_envs, self.num_agents, self.npc_capacity
        shapes = {
            "obs": (E, N, _lib.OBS_DIM), "reward": (E, N), "done": (E, N), "status": (E, N), "terminated": (E,),
            "truncated": (E,), "agents_alive": (E,), "step": (E,), "lidar_hit": (E, N, _lib.MAX_RAYS),
            "npc_count": (E,), "tick": (E,),
        }
        self.buf: Dict[str, torch.Tensor] = {}
        for name, ts in _lib._BUF_FIELDS:
            ptr = getattr(b, name)
            if name == "events":
                shape, t = (E, 6), "i4"
            else:
                t = _TORCH_VIEW.get(ts, ts)
                if name in shapes:
                    shape = shapes[name]
                elif name.startswith("ego_"):
                    shape = (E, N)
                else:
                    shape = (E, M)
            with torch.cuda.device(self.device):
                self.buf[name] = torch.as_tensor(_DevArray(ptr, shape, t, self), device=self.device)

    def _stream(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ API
    def reset(self, mask: Optional[torch.Tensor] = None):
        """reset() + add_car_with_route for every env (or those where mask != 0).  Returns (obs, {})."""
        mp = None
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            if mask.numel() != self.num_envs:
                raise ValueError("mask must have num_envs elements")
            mp = C.c_void_p(mask.data_ptr())
        _lib.check(self._lib, self._lib.isx_reset(self._h, mp, self._stream()))
        return self.buf["obs"], {}

    def step(self, actions: torch.Tensor, dt: float = 1.0 / 60.0):
        if not isinstance(actions, torch.Tensor):
            actions = torch.as_tensor(np.asarray(actions, dtype=np.float32))
        a = actions.to(device=self.device, dtype=torch.float32).contiguous()
        if a.numel() != self.num_envs * self.num_agents * 2:
            raise ValueError(f"Expected actions shape ({self.num_envs},{self.num_agents},2), got {tuple(actions.shape)}")
        _lib.check(self._lib, self._lib.isx_step(self._h, C.c_void_p(a.data_ptr()), C.c_float(dt), self._stream()))
        b = self.buf
        info = {"done": b["done"], "status": b["status"], "agents_alive": b["agents_alive"], "step": b["step"],
                "lidar_hit": b["lidar_hit"], "npc_count": b["npc_count"], "events": b["events"]}
        return b["obs"], b["reward"], b["terminated"].bool(), b["truncated"].bool(), info

    def _host_views(self):
        if getattr(self, "_hv", None) is None:
            ptrs = [C.c_void_p() for _ in range(7)]
            _lib.check(self._lib, self._lib.isx_host_views(self._h, *[C.byref(p) for p in ptrs]))
            E, N = self.num_envs, self.num_agents

            def view(p, shape, ctype, dtype):
                n = int(np.prod(shape))
                return np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), shape=(n,)).view(dtype).reshape(shape)

            self._hv = dict(actions=view(ptrs[0], (E, N, 2), C.c_float, np.float32), obs=view(ptrs[1], (E, N, _lib.OBS_DIM), C.c_float, np.float32),
                            reward=view(ptrs[2], (E, N), C.c_float, np.float32), done=view(ptrs[3], (E, N), C.c_uint8, np.uint8),
                            status=view(ptrs[4], (E, N), C.c_uint8, np.uint8), terminated=view(ptrs[5], (E,), C.c_uint8, np.uint8),
                            truncated=view(ptrs[6], (E,), C.c_uint8, np.uint8))
            aux = [C.c_void_p(), C.c_void_p()]
            _lib.check(self._lib, self._lib.isx_host_views_aux(self._h, C.byref(aux[0]), C.byref(aux[1])))
            self._hv["agents_alive"] = view(aux[0], (E,), C.c_int32, np.int32)
            self._hv["step"] = view(aux[1], (E,), C.c_int32, np.int32)
        return self._hv

    @property
    def host_actions(self) -> np.ndarray:
        """The pinned [E, N, 2] float32 staging buffer the host step uploads from: a policy that writes its output here in
        place (and then calls ``step_host(None)``) saves the staging copy (0.3 ms per step at 65,536 envs x 8 agents)."""
        return self._host_views()["actions"]

    def step_host(self, actions: Optional[np.ndarray], dt: float = 1.0 / 60.0, copy: bool = False):
        """The same step through HOST buffers (what env.py's list<->numpy conversions amount to): host actions in, host
        obs / reward / done / status / terminated / truncated out.  The returned arrays are views of the library's pinned
        staging buffers, overwritten by the next host step (copy=True detaches them).  Over PCIe travels a compact record per
        agent (31 features + lidar hit indices); the library's host threads rebuild the float32 rows bit-identically while
        later env ranges are still being simulated and copied (isx_host_step_info / host_step_bytes())."""
        hv = self._host_views()
        if actions is None:                            # the caller already wrote this step's actions into host_actions
            _lib.check(self._lib, self._lib.isx_step_pinned(self._h, C.c_float(dt), self._stream()))
        else:
            a = np.ascontiguousarray(actions, dtype=np.float32)
            if a.size != hv["actions"].size:
                raise ValueError(f"Expected actions shape {hv['actions'].shape}, got {a.shape}")
            # isx_step_host stages the actions with the library's host threads; NULL outputs = results stay in the pinned views
            _lib.check(self._lib, self._lib.isx_step_host(self._h, C.c_void_p(a.ctypes.data), C.c_float(dt), None, None, None, None, None, None,
                                                         self._stream()))
        out = (hv["obs"], hv["reward"], hv["done"], hv["status"], hv["terminated"].astype(bool), hv["truncated"].astype(bool))
        if copy:
            out = tuple(np.array(x) for x in out)
        return out

    def check_guards(self) -> int:
        """Debug aid (ISX_GUARD=1 in the environment when the batch was created): number of red zones around the library's
        device buffers that a kernel wrote into; 0 = no out-of-bounds store since creation (isx_debug_check_guards)."""
        v = C.c_int64(-1)
        _lib.check(self._lib, self._lib.isx_debug_check_guards(self._h, C.byref(v)))
        return int(v.value)

    def host_step_bytes(self) -> Dict[str, Any]:
        """What one step_host() moves over PCIe and who completes the obs rows on the host (isx_host_step_info)."""
        a, b, t, r = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32()
        _lib.check(self._lib, self._lib.isx_host_step_info(self._h, C.byref(a), C.byref(b), C.byref(t), C.byref(r)))
        return {"h2d": a.value, "d2h": b.value, "host_expand_threads": t.value, "pipeline_ranges": r.value,
                "obs_transport": ("compact: 32 f32 + lidar_rays u8 per agent over PCIe, 127-float rows rebuilt bit-identically by host threads"
                                  if t.value > 0 else "rows: the 127-float rows themselves cross PCIe into the pinned view (small batch)")}

    def rollout(self, steps: int, dt: float = 1.0 / 60.0):
        """`steps` steps with on-device Philox actions (random-action rollout of BASELINE.json)."""
        _lib.check(self._lib, self._lib.isx_rollout(self._h, int(steps), C.c_float(dt), self._stream()))

    def rollout_timed(self, steps: int, dt: float = 1.0 / 60.0):
        """rollout() with per-kernel CUDA-event timing; returns (ms in k_dynamics, ms in k_lidar_obs)."""
        a, b = C.c_float(), C.c_float()
        _lib.check(self._lib, self._lib.isx_rollout_timed(self._h, int(steps), C.c_float(dt), self._stream(), C.byref(a), C.byref(b)))
        return a.value, b.value

    def rollout_timed4(self, steps: int, dt: float = 1.0 / 60.0):
        """per-kernel CUDA-event timing: (ms k_traffic, ms k_ego, ms k_features, ms k_lidar_obs) summed over `steps`."""
        ms = (C.c_float * 4)()
        _lib.check(self._lib, self._lib.isx_rollout_timed4(self._h, int(steps), C.c_float(dt), self._stream(), ms))
        return tuple(ms)

    # ------------------------------------------------------------------ run-time settings (no buffers are rebuilt)
    def set_lidar_rays(self, rays: int):
        """Beam count of every lidar (Lidar.rays, Lidar.h:11): clears the stored hits and refreshes obs."""
        _lib.check(self._lib, self._lib.isx_set_lidar_rays(self._h, int(rays)))
        self.lidar_rays = int(rays)
        for gc in self.group_configs:
            gc["lidar_rays"] = int(rays)

    def set_reward_config(self, reward_cfg=None, group: int = -1):
        """reward_config.* of the reference env object (bindings.cpp:33-42); dict with env.py's keys or an 8-vector."""
        vec = reward_vector(reward_cfg) if (reward_cfg is None or isinstance(reward_cfg, dict)) else tuple(float(x) for x in reward_cfg)
        _lib.check(self._lib, self._lib.isx_set_reward(self._h, int(group), (C.c_float * 8)(*vec)))

    def configure(self, use_team: bool, respawn: bool, max_steps: int, group: int = -1):
        """IntersectionEnv.configure(use_team, respawn, max_steps) (IntersectionEnv.cpp:50-54) on the live batch."""
        _lib.check(self._lib, self._lib.isx_configure_episode(self._h, int(group), int(bool(use_team)), int(bool(respawn)), int(max_steps)))
        if group < 0:
            self.max_steps = int(max_steps)

    def set_traffic_density(self, density: float, group: int = -1):
        _lib.check(self._lib, self._lib.isx_set_traffic_density(self._h, int(group), C.c_float(max(0.0, float(density)))))

    def observe(self):
        _lib.check(self._lib, self._lib.isx_observe(self._h, self._stream()))
        return self.buf["obs"]

    def render(self, env: int = 0) -> torch.Tensor:
        """Headless debug picture of one env: uint8 cuda tensor [750, 750, 3] (road, cars, the lidar beams that hit)."""
        img = torch.empty((750, 750, 3), dtype=torch.uint8, device=self.device)
        _lib.check(self._lib, self._lib.isx_render(self._h, int(env), C.c_void_p(img.data_ptr()), self._stream()))
        return img

    # ------------------------------------------------------------------ snapshots (get_state / set_state for the batch)
    def snapshot(self):
        """Save the full state of every env on the device (EnvState of the reference, for MCTS-style rollbacks)."""
        s = C.c_void_p()
        _lib.check(self._lib, self._lib.isx_snapshot_create(self._h, C.byref(s)))
        _lib.check(self._lib, self._lib.isx_snapshot_save(self._h, s, self._stream()))
        self._snaps = getattr(self, "_snaps", [])
        self._snaps.append(s)
        return s

    def save_into(self, snap):
        _lib.check(self._lib, self._lib.isx_snapshot_save(self._h, snap, self._stream()))

    def restore(self, snap, mask: Optional[torch.Tensor] = None):
        """Roll every env (or those where mask != 0) back to the snapshot."""
        mp = None
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            mp = C.c_void_p(mask.data_ptr())
        _lib.check(self._lib, self._lib.isx_snapshot_restore(self._h, snap, mp, self._stream()))

    def stats(self) -> Dict[str, Any]:
        s = _lib.Stats()
        torch.cuda.synchronize(self.device)
        _lib.check(self._lib, self._lib.isx_stats_read(self._h, C.byref(s)))
        return {"agent_steps": s.agent_steps, "status_hist": {STATUS_NAMES[i]: s.status_hist[i] for i in range(6)},
                "npc_spawned": s.npc_spawned, "npc_removed": s.npc_removed, "npc_collided": s.npc_collided,
                "npc_overflow": s.npc_overflow, "env_resets": s.env_resets, "reward_sum": s.reward_sum,
                "neighbor_tie_sorts": s.neighbor_tie_sorts}

    def stats_tensors(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Device views for the one collective of this path: (int64[15] counters, float64[1] reward_sum).  They are two
        tensors on purpose — all-reduce(sum) the counters as integers and reward_sum as a double (`reduce_stats` does)."""
        p, r, n = C.c_void_p(), C.c_void_p(), C.c_int32()
        _lib.check(self._lib, self._lib.isx_stats_device_ptrs(self._h, C.byref(p), C.byref(n), C.byref(r), self._stream()))
        return (torch.as_tensor(_DevArray(p.value, (n.value,), "i8", self), device=self.device),
                torch.as_tensor(_DevArray(r.value, (1,), "f8", self), device=self.device))

    def reduce_stats(self, group=None) -> Dict[str, Any]:
        """Job totals of the episode counters over the ranks of `group`: the only collective the path has."""
        cnt, rs = self.stats_tensors()
        return reduce_stat_tensors(cnt.clone(), rs.clone(), group)

    def reset_stats(self):
        _lib.check(self._lib, self._lib.isx_stats_reset(self._h))

    def get_env_state(self, env: int):
        N, M = self.num_agents, self.npc_capacity
        egos = (_lib.CarState * N)()
        npcs = (_lib.CarState * M)()
        n, sc, tk = C.c_int32(), C.c_int32(), C.c_uint32()
        _lib.check(self._lib, self._lib.isx_get_env_state(self._h, env, egos, npcs, M, C.byref(n), C.byref(sc), C.byref(tk)))
        return egos, npcs, n.value, sc.value, tk.value

    def set_env_state(self, env: int, egos, npcs, n_npcs: int, step_count: int, tick: int):
        _lib.check(self._lib, self._lib.isx_set_env_state(self._h, env, egos, npcs, int(n_npcs), int(step_count), int(tick)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.buf = {}
            self._hv = None
            for sn in getattr(self, "_snaps", []):
                self._lib.isx_snapshot_destroy(sn)
            self._snaps = []
            self._lib.isx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
