"""BASELINE-length parity census, in the driver-run suite: every BASELINE.json config shape (C1..C5; C4 with the default
96-beam lidar, C5 with 72) x 2000 steps x 64 envs x seeds {0, 1, 2}, free-running (no state injection), against the
reference's own C++ (oracle/_ref) — or the pinned C port where _ref did not travel.  Compared bit for bit EVERY step:
obs, rewards, done / status, terminated / truncated, agents_alive, step, lidar hit indices, NPC spawn / removal events,
NPC count and poses.  (north_star: "2000-step rollout ... identical seeds, actions and routes".)

The three seeds run as three config groups of one heterogeneous batch (isx_create_groups), 192 envs per config; the CPU
checkers step on all host cores (ctypes releases the GIL)."""
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
def test_census_2000_steps_every_bit(name):
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
    pool = ThreadPoolExecutor(os.cpu_count() or 4)
    obs0, _ = b.reset()
    torch.cuda.synchronize()
    assert (u32(obs0.cpu().numpy()) == u32(np.stack([r.obs() for r in refs]))).all()
    hist = np.zeros(6, np.int64)
    spawned = removed = 0

    def one(args):
        e, act, explicit = args
        r = refs[e]
        o = r.step(act)
        ev = r.events() if traffic else None
        npc = r.npcs() if traffic else None
        lid = [r.lidar(a) for a in range(N)] if explicit else None
        return o, ev, npc, lid

    for t in range(STEPS):
        act = np.stack([po.philox_actions(seed_of[e], id_of[e], refs[e].tick + 1, N) for e in range(E)])
        b.step(torch.from_numpy(act).cuda())
        explicit = (t % 40 == 0)
        outs = list(pool.map(one, [(e, act[e], explicit) for e in range(E)]))
        torch.cuda.synchronize()
        g = {k: b.buf[k].cpu().numpy() for k in ("obs", "reward", "done", "status", "terminated", "truncated", "agents_alive", "step", "lidar_hit")}
        want_obs = np.stack([o["obs"] for o, _, _, _ in outs])
        bad = u32(g["obs"]) != u32(want_obs)
        assert not bad.any(), (name, "obs", t, np.argwhere(bad)[:3].tolist())
        assert (u32(g["reward"]) == u32(np.stack([o["reward"] for o, _, _, _ in outs]))).all(), (name, "reward", t)
        assert (g["done"] == np.stack([o["done"] for o, _, _, _ in outs])).all() and (g["status"] == np.stack([o["status"] for o, _, _, _ in outs])).all(), (name, "status", t)
        term = np.array([o["terminated"] for o, _, _, _ in outs])
        trunc = np.array([o["truncated"] for o, _, _, _ in outs])
        assert (g["terminated"].astype(bool) == term).all() and (g["truncated"].astype(bool) == trunc).all(), (name, "flags", t)
        assert (g["agents_alive"] == np.array([o["agents_alive"] for o, _, _, _ in outs])).all() and (g["step"] == np.array([o["step"] for o, _, _, _ in outs])).all()
        # lidar hit indices: obs[31+i] = float(4k) * (1/250) is injective in k, so the reference's k is read back from its obs
        # (all egos stay alive in these configs); every 40th step the reference's distance vectors are compared directly as well
        d_ref = want_obs[:, :, 31:31 + R].astype(np.float64) * 250.0          # 4k for a hit at sample k <= 62, 250 for none
        k_ref = np.where(d_ref > 249.0, 0, np.rint(d_ref / 4.0)).astype(np.int64)
        assert (g["lidar_hit"][:, :, :R].astype(np.int64) == k_ref).all(), (name, "lidar hit index", t)
        if explicit:
            for e, (_, _, _, lid) in enumerate(outs):
                for a in range(N):
                    k = np.where(lid[a] >= 250.0, 0, lid[a] / 4.0).astype(np.int64)
                    assert (g["lidar_hit"][e, a, :R].astype(np.int64) == k).all(), (name, "lidar distances", t, e, a)
        if traffic:
            ev = b.buf["events"].cpu().numpy()
            nc = b.buf["npc_count"].cpu().numpy()
            want_ev = np.array([[int(v["rng_draws"]), int(v["spawn_route"]), int(v["spawned"]), int(v["removed_mask"]), int(v["collided_mask"]), int(v["npc_count"])]
                                for _, v, _, _ in outs], np.int64)
            got_ev = ev.astype(np.int64)
            got_ev[:, 3:5] &= 0xFFFFFFFF
            cols = [0, 1, 2, 3, 5] if cls is po.RefEnv else [0, 1, 2, 3, 4, 5]      # the reference driver does not infer collided_mask
            assert (got_ev[:, cols] == want_ev[:, cols]).all(), (name, "events", t)
            assert (nc == want_ev[:, 5]).all(), (name, "npc_count", t)
            spawned += int(want_ev[:, 2].sum())
            removed += int(sum(bin(int(m)).count("1") for m in want_ev[:, 3]))
            pose = {k: b.buf[k].cpu().numpy() for k in ("npc_x", "npc_y", "npc_v", "npc_heading")}
            for e, (_, _, npc, _) in enumerate(outs):
                n = len(npc)
                for f, k in (("x", "npc_x"), ("y", "npc_y"), ("v", "npc_v"), ("heading", "npc_heading")):
                    assert (u32(pose[k][e, :n]) == u32(npc[f])).all(), (name, "npc " + f, t, e)
        for o, _, _, _ in outs:
            hist += np.bincount(o["status"], minlength=6)
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
    pool.shutdown()
    b.close()
