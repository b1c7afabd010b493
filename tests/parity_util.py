"""Shared driver for the GPU parity tests: steps the CUDA stepper (through the C ABI) and E independent CPU
checker envs (oracle/_ref = the reference's own C++ when present, else the C restatement) on the same
Philox action stream, and compares every output and every piece of state, bit for bit."""
from __future__ import annotations

import numpy as np

import pyoracle as po


def checker_class():
    """The reference build itself when it travelled with the repo, else the C restatement."""
    return po.RefEnv if po.have_ref() else po.OracleEnv


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


class Mismatch(AssertionError):
    pass


def make_pair(benv_cls, cfg, checker=None, seed=0, env_id_base=0):
    """cfg: dict for BatchedIntersectionEnv.  Returns (cuda_env, [checker envs])."""
    checker = checker or checker_class()
    c = dict(cfg)
    c["seed"] = seed
    c["env_id_base"] = env_id_base
    b = benv_cls(c)
    refs = []
    for e in range(b.num_envs):
        refs.append(checker(num_lanes=b.num_lanes, ego_routes=b.ego_routes, use_team=bool(cfg.get("use_team_reward", False)),
                            respawn=bool(cfg.get("respawn_enabled", True)), max_steps=int(cfg.get("max_steps", 2000)),
                            traffic=bool(cfg.get("traffic_flow", False)), density=float(cfg.get("traffic_density", 0.5)),
                            traffic_routes=b.traffic_routes, reward=tuple(_reward_vec(cfg)), lidar_rays=b.lidar_rays,
                            seed=seed, env_id=env_id_base + e))
    return b, refs


def make_group_pair(benv_cls, cfgs, checker=None):
    """Heterogeneous batch (list of config dicts, each with its own seed / env_id_base) and one checker env per env,
    configured like the group the env belongs to."""
    checker = checker or checker_class()
    b = benv_cls([dict(c) for c in cfgs])
    refs = []
    for cfg, gc in zip(cfgs, b.group_configs):
        for e in range(gc["num_envs"]):
            refs.append(checker(num_lanes=gc["num_lanes"], ego_routes=gc["ego_routes"], use_team=bool(cfg.get("use_team_reward", False)),
                                respawn=bool(cfg.get("respawn_enabled", True)), max_steps=int(cfg.get("max_steps", 2000)),
                                traffic=bool(cfg.get("traffic_flow", False)), density=float(cfg.get("traffic_density", 0.5)),
                                traffic_routes=gc["traffic_routes"], reward=tuple(_reward_vec(cfg)), lidar_rays=gc["lidar_rays"],
                                seed=int(cfg.get("seed", 0)), env_id=int(cfg.get("env_id_base", 0)) + e))
    return b, refs


def _reward_vec(cfg):
    from marl_traffic_intersection_b200.utils import reward_vector
    return reward_vector(cfg.get("reward_config"))


def compare_step(b, refs, outs, where, check_state=True):
    """b: BatchedIntersectionEnv after a step; refs: checker envs after the same step; outs: their step dicts."""
    import torch
    torch.cuda.synchronize()
    buf = {k: v.cpu().numpy() for k, v in b.buf.items()}
    E, N = b.num_envs, b.num_agents
    for e in range(E):
        o = outs[e]
        tag = f"{where} env {e}"
        if not (bits(buf["obs"][e]) == bits(o["obs"])).all():
            idx = np.argwhere(bits(buf["obs"][e]) != bits(o["obs"]))
            a, k = idx[0]
            raise Mismatch(f"{tag}: obs[{a},{k}] cuda={buf['obs'][e][a, k]!r} ref={o['obs'][a, k]!r} ({len(idx)} diffs)")
        if not (bits(buf["reward"][e]) == bits(o["reward"])).all():
            raise Mismatch(f"{tag}: reward cuda={buf['reward'][e]} ref={o['reward']}")
        if not (buf["done"][e] == o["done"]).all() or not (buf["status"][e] == o["status"]).all():
            raise Mismatch(f"{tag}: done/status cuda={buf['done'][e]},{buf['status'][e]} ref={o['done']},{o['status']}")
        if bool(buf["terminated"][e]) != o["terminated"] or bool(buf["truncated"][e]) != o["truncated"]:
            raise Mismatch(f"{tag}: terminated/truncated")
        if int(buf["agents_alive"][e]) != o["agents_alive"] or int(buf["step"][e]) != o["step"]:
            raise Mismatch(f"{tag}: agents_alive/step {buf['agents_alive'][e]} {buf['step'][e]} vs {o['agents_alive']} {o['step']}")
        if not check_state:
            continue
        r = refs[e]
        eg = r.egos()
        for f, name in (("x", "ego_x"), ("y", "ego_y"), ("v", "ego_v"), ("heading", "ego_heading"), ("steer", "ego_steer"),
                        ("acc", "ego_acc"), ("prev_dist", "ego_prev_dist"), ("prev_a0", "ego_prev_a0"), ("prev_a1", "ego_prev_a1")):
            if not (bits(buf[name][e]) == bits(eg[f])).all():
                raise Mismatch(f"{tag}: ego {f} cuda={buf[name][e]} ref={eg[f]}")
        if not (buf["ego_path_index"][e] == eg["path_index"]).all():
            raise Mismatch(f"{tag}: ego path_index")
        # lidar hit indices (bit-exact): distance = 4k, k in 1..62, or 250 = none
        for a in range(N):
            d = r.lidar(a)
            k = np.where(d >= 250.0, 0, (d / 4.0)).astype(np.int64)
            got = buf["lidar_hit"][e, a, : len(d)].astype(np.int64)
            alive = eg["alive"][a]
            if alive and not (got == k).all():
                i = int(np.argwhere(got != k)[0][0])
                raise Mismatch(f"{tag}: lidar hit index agent {a} beam {i}: cuda={got[i]} ref={k[i]}")
        if b.traffic_flow:
            ev = r.events()
            ce = buf["events"][e]
            got = dict(rng_draws=int(ce[0]), spawn_route=int(ce[1]), spawned=int(ce[2]), removed_mask=int(ce[3]) & 0xFFFFFFFF,
                       npc_count=int(ce[5]))
            for f in got:
                if got[f] != int(ev[f]):
                    raise Mismatch(f"{tag}: event {f}: cuda={got[f]} ref={int(ev[f])} (all: cuda={got} ref={ev})")
            npc = r.npcs()
            n = len(npc)
            if int(buf["npc_count"][e]) != n:
                raise Mismatch(f"{tag}: npc_count {buf['npc_count'][e]} vs {n}")
            for f, name in (("x", "npc_x"), ("y", "npc_y"), ("v", "npc_v"), ("heading", "npc_heading"), ("steer", "npc_steer")):
                if not (bits(buf[name][e, :n]) == bits(npc[f])).all():
                    raise Mismatch(f"{tag}: npc {f} cuda={buf[name][e, :n]} ref={npc[f]}")
            if not (buf["npc_path_index"][e, :n] == npc["path_index"]).all() or not (buf["npc_route"][e, :n] == npc["route"]).all():
                raise Mismatch(f"{tag}: npc path_index/route")
            if not (buf["npc_uid"][e, :n].astype(np.int64) & 0xFFFFFFFF == npc["uid"].astype(np.int64)).all():
                raise Mismatch(f"{tag}: npc uid {buf['npc_uid'][e, :n]} vs {npc['uid']}")
    return buf


def free_run(b, refs, steps, seed, env_id_base=0, policy=None, check_state=True, dt=1.0 / 60.0, host_api=False):
    """Free-running parity: both sides evolve on their own; no state is injected.  policy(obs[E,N,127], t) -> actions
    or None for the Philox random stream.  Envs that terminate / truncate are reset on both sides (what a user does)."""
    import torch
    E, N = b.num_envs, b.num_agents
    obs0, _ = b.reset()
    torch.cuda.synchronize()
    o0 = obs0.cpu().numpy()
    for e in range(E):
        refs[e].reset()
        if not (bits(o0[e]) == bits(refs[e].obs())).all():
            raise Mismatch(f"reset obs env {e}")
    cur_obs = o0
    counts = np.zeros(6, np.int64)
    for t in range(steps):
        if policy is None:
            act = np.stack([po.philox_actions(seed, env_id_base + e, refs[e].tick + 1, N) for e in range(E)])
        else:
            act = np.asarray(policy(cur_obs, t), np.float32).reshape(E, N, 2)
        if host_api:
            b.step_host(act, dt)
        else:
            b.step(torch.from_numpy(act).cuda(), dt)
        outs = [refs[e].step(act[e], dt) for e in range(E)]
        buf = compare_step(b, refs, outs, f"step {t + 1}", check_state=check_state)
        for e in range(E):
            for s in outs[e]["status"]:
                counts[s] += 1
        cur_obs = buf["obs"].copy()
        need = np.array([o["terminated"] or o["truncated"] for o in outs])
        if need.any():
            b.reset(torch.from_numpy(need.astype(np.uint8)).cuda())
            torch.cuda.synchronize()
            fresh = b.buf["obs"].cpu().numpy()
            for e in np.nonzero(need)[0]:
                refs[e].reset()
                if not (bits(fresh[e]) == bits(refs[e].obs())).all():
                    raise Mismatch(f"obs after reset, env {e}, step {t + 1}")
                cur_obs[e] = fresh[e]
    return counts


def pursuit_policy(throttle=0.35, gain=2.0):
    """Follows the route using obs[5] (heading error to the look-ahead point): reaches SUCCESS, so the success /
    terminated branches and long NPC interactions get exercised."""
    def pol(obs, t):
        steer = np.clip(gain * obs[..., 5], -1.0, 1.0)
        thr = np.where(obs[..., 2] * 8.0 < 4.0, throttle, 0.0)
        return np.stack([thr, steer], axis=-1).astype(np.float32)
    return pol
