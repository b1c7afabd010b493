"""BASELINE-length parity census, in the driver-run suite: every BASELINE.json config shape (C1..C5; C4 with the default
96-beam lidar, C5 with 72) x 2000 steps x 64 envs x seeds {0, 1, 2}, free-running (no state injection), against the
reference's own C++ (oracle/_ref) — or the pinned C port where _ref did not travel.  Compared bit for bit EVERY step:
obs, rewards, done / status, terminated / truncated, agents_alive, step, lidar hit indices, NPC spawn / removal events,
NPC count and poses.  (north_star: "2000-step rollout ... identical seeds, actions and routes".)

The three seeds run as three config groups of one heterogeneous batch (isx_create_groups), 192 envs per config; the CPU
checkers step on all host cores (isxref_step_batch: one library call per step for all envs)."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import pyoracle as po
from parity_util import checker_class

pytestmark = pytest.mark.gpu

R3 = po.ROUTES_3LANES
ENVS, STEPS, SEEDS = 64, 2000, (0, 1, 2)
CENSUS = {
    "C1_single": dict(num_agents=1, num_lanes=3, ego_routes=[("IN_6", "OUT_2")]),
    "C2_team3": dict(num_agents=3, num_lanes=3, ego_routes=[("IN_6", "OUT_2"), ("IN_4", "OUT_8"), ("IN_5", "OUT_7")], use_team_reward=True),
    "C3_traffic": dict(num_agents=1, num_lanes=3, ego_routes=[("IN_6", "OUT_2")], traffic_flow=True, traffic_density=0.5),
    "C4_eight_96beams": dict(num_agents=8, num_lanes=3, ego_routes=R3[:8], lidar_rays=96),
    "C5_eight_traffic_72beams": dict(num_agents=8, num_lanes=3, ego_routes=R3[:8], traffic_flow=True, traffic_density=1.0, lidar_rays=72),
}


def u32(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("name", list(CENSUS))
def test_census_2000_steps_every_bit(name, monkeypatch):
    if name.startswith("C5"):          # the k_traffic instance the 65,536-env configuration runs (a 192-env batch would get 32 lanes)
        monkeypatch.setenv("ISX_TRAFFIC_LANES", "8")
    import torch
    from marl_traffic_intersection_b200 import BatchedIntersectionEnv
    base = CENSUS[name]
    cfgs = [dict(base, num_envs=ENVS, seed=s, env_id_base=0, max_steps=2000) for s in SEEDS]
    b = BatchedIntersectionEnv(cfgs)
    E, N, R, M = b.num_envs, b.num_agents, b.lidar_rays, b.npc_capacity
    traffic = b.traffic_flow
    cls = checker_class()
    refs, seed_of, id_of = [], [], []
    for s in SEEDS:
        for e in range(ENVS):
            refs.append(cls(num_lanes=3, ego_routes=base["ego_routes"], use_team=bool(base.get("use_team_reward", False)), traffic=traffic,
                            density=float(base.get("traffic_density", 0.5)), lidar_rays=R, seed=s, env_id=e, max_steps=2000))
            seed_of.append(s)
            id_of.append(e)
    obs0, _ = b.reset()
    torch.cuda.synchronize()
    assert (u32(obs0.cpu().numpy()) == u32(np.stack([r.obs() for r in refs]))).all()
    hist = np.zeros(6, np.int64)
    spawned = removed = 0
    batch = cls is po.RefEnv                                   # one library call per step for all envs (isxref_step_batch)
    pool = None if batch else ThreadPoolExecutor(os.cpu_count() or 4)

    def slow_step(act):
        """per-env ctypes path (the C port has no batch entry point): same dict as pyoracle.step_batch"""
        def one(e):
            r = refs[e]
            o = r.step(act[e])
            lk = np.zeros((N, 96), np.uint8)
            for a in range(N):
                d = r.lidar(a)
                lk[a, :len(d)] = np.where(d >= 250.0, 0, d / 4.0).astype(np.uint8)
            pose = np.zeros((M, 4), np.float32)
            ev = r.events() if traffic else np.zeros(1, po.EVENTS_DTYPE)[0]
            if traffic:
                npc = r.npcs()
                for j, f in enumerate(("x", "y", "v", "heading")):
                    pose[:len(npc), j] = npc[f]
            return o, lk, pose, ev
        res = list(pool.map(one, range(E)))
        out = {k: np.stack([o[k] for o, _, _, _ in res]) for k in ("obs", "reward", "done", "status")}
        for k in ("terminated", "truncated", "agents_alive", "step"):
            out[k] = np.array([int(o[k]) for o, _, _, _ in res])
        out["lidar_k"] = np.stack([x[1] for x in res])
        out["npc_pose"] = np.stack([x[2] for x in res])
        out["events"] = np.array([x[3] for x in res], po.EVENTS_DTYPE)
        return out

    ticks = np.zeros(E, np.int64)
    for t in range(STEPS):
        ticks += 1
        act = po.philox_actions_batch(seed_of, id_of, ticks, N)
        b.step(torch.from_numpy(act).cuda())
        o = po.step_batch(refs, act, lidar=True, npc_cap=M if traffic else 0) if batch else slow_step(act)
        torch.cuda.synchronize()
        g = {k: b.buf[k].cpu().numpy() for k in ("obs", "reward", "done", "status", "terminated", "truncated", "agents_alive", "step", "lidar_hit")}
        bad = u32(g["obs"]) != u32(o["obs"])
        assert not bad.any(), (name, "obs", t, np.argwhere(bad)[:3].tolist())
        assert (u32(g["reward"]) == u32(o["reward"])).all(), (name, "reward", t)
        assert (g["done"] == o["done"]).all() and (g["status"] == o["status"]).all(), (name, "status", t)
        term, trunc = o["terminated"].astype(bool), o["truncated"].astype(bool)
        assert (g["terminated"].astype(bool) == term).all() and (g["truncated"].astype(bool) == trunc).all(), (name, "flags", t)
        assert (g["agents_alive"] == o["agents_alive"]).all() and (g["step"] == o["step"]).all(), (name, "agents_alive/step", t)
        # lidar hit indices, read from the reference's Lidar::distances of every ego (all egos stay alive in these configs)
        assert (g["lidar_hit"][:, :, :R] == o["lidar_k"][:, :, :R]).all(), (name, "lidar hit index", t)
        if traffic:
            ev = b.buf["events"].cpu().numpy().astype(np.int64)
            ev[:, 3:5] &= 0xFFFFFFFF
            want = o["events"]
            for col, f in ((0, "rng_draws"), (1, "spawn_route"), (2, "spawned"), (3, "removed_mask"), (5, "npc_count")):
                assert (ev[:, col] == want[f].astype(np.int64)).all(), (name, "event " + f, t)
            if not batch:                                       # the reference driver does not infer collided_mask
                assert (ev[:, 4] == want["collided_mask"].astype(np.int64)).all(), (name, "event collided_mask", t)
            nc = b.buf["npc_count"].cpu().numpy()
            assert (nc == want["npc_count"]).all(), (name, "npc_count", t)
            spawned += int(want["spawned"].sum())
            removed += int(sum(bin(int(m)).count("1") for m in want["removed_mask"]))
            live = np.arange(M)[None, :] < nc[:, None]
            for j, k in enumerate(("npc_x", "npc_y", "npc_v", "npc_heading")):
                got = b.buf[k].cpu().numpy()
                assert (u32(got)[live] == u32(o["npc_pose"][:, :, j])[live]).all(), (name, k, t)
        hist += np.bincount(o["status"].reshape(-1), minlength=6)
        need = term | trunc
        if need.any():
            b.reset(torch.from_numpy(need.astype(np.uint8)).cuda())
            torch.cuda.synchronize()
            fresh = b.buf["obs"].cpu().numpy()
            for e in np.nonzero(need)[0]:
                refs[e].reset()
                assert (u32(fresh[e]) == u32(refs[e].obs())).all(), (name, "obs after reset", t, e)
    assert hist.sum() == E * N * STEPS and hist[po.STATUS_NAMES.index("ALIVE")] > 0.9 * hist.sum()
    assert hist[3] + hist[4] + hist[5] > 0                    # crashes (wall / line / car) were exercised
    if traffic:
        assert spawned > 100 and removed > 50
    print(f"census {name}: {E} envs x {N} agents x {STEPS} steps = {E * N * STEPS} agent-steps, checker {cls.__name__}, "
          f"0 mismatches, status histogram {dict(zip(po.STATUS_NAMES, hist.tolist()))}, npc spawned/removed {spawned}/{removed}")
    if pool is not None:
        pool.shutdown()
    b.close()
