#!/usr/bin/env python
"""Large free-running parity census (one-off evidence, slower than the test-suite cases): the CUDA stepper against the
reference's own C++ (oracle/_ref) — or the C port where _ref is absent — on BASELINE.json config shapes, 2000 steps,
hundreds of envs, every obs / reward / status / flag / lidar hit index / NPC event compared bit for bit each step.
The checker envs run on all host cores (threads; ctypes releases the GIL).

    python tools/parity_census.py [config ...]        # configs: C2 C3 C4 C5 (default all)
Prints one JSON line per config; exit code 1 on any mismatch."""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import pyoracle as po  # noqa: E402
from marl_traffic_intersection_b200 import BatchedIntersectionEnv  # noqa: E402
from parity_util import checker_class, make_pair  # noqa: E402

R3 = po.ROUTES_3LANES
CONFIGS = {
    "C2": (dict(num_envs=256, num_agents=3, num_lanes=3, ego_routes=[("IN_6", "OUT_2"), ("IN_4", "OUT_8"), ("IN_5", "OUT_7")], use_team_reward=True), 2000),
    "C3": (dict(num_envs=256, num_agents=1, num_lanes=3, ego_routes=[("IN_6", "OUT_2")], traffic_flow=True, traffic_density=0.5), 2000),
    "C4": (dict(num_envs=96, num_agents=8, num_lanes=3, ego_routes=R3[:8]), 2000),
    "C5": (dict(num_envs=96, num_agents=8, num_lanes=3, ego_routes=R3[:8], traffic_flow=True, traffic_density=1.0, lidar_rays=72), 2000),
}


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def run(name, seed=0):
    cfg, steps = CONFIGS[name]
    b, refs = make_pair(BatchedIntersectionEnv, cfg, seed=seed)
    E, N = b.num_envs, b.num_agents
    pool = ThreadPoolExecutor(os.cpu_count() or 4)
    b.reset()
    for r in refs:
        r.reset()
    mism = dict(obs=0, reward=0, status=0, flags=0, lidar=0, events=0, npc=0)
    hist = np.zeros(6, np.int64)
    t0 = time.time()
    for t in range(steps):
        act = np.stack([po.philox_actions(seed, e, refs[e].tick + 1, N) for e in range(E)])
        b.step(torch.from_numpy(act).cuda())

        def one(e):
            o = refs[e].step(act[e])
            return o, [refs[e].lidar(a) for a in range(N)], (refs[e].events() if b.traffic_flow else None), (len(refs[e].npcs()) if b.traffic_flow else 0)

        outs = list(pool.map(one, range(E)))
        torch.cuda.synchronize()
        obs = b.buf["obs"].cpu().numpy(); rew = b.buf["reward"].cpu().numpy(); st = b.buf["status"].cpu().numpy()
        term = b.buf["terminated"].cpu().numpy(); trunc = b.buf["truncated"].cpu().numpy(); hit = b.buf["lidar_hit"].cpu().numpy()
        ev = b.buf["events"].cpu().numpy(); nc = b.buf["npc_count"].cpu().numpy()
        need = np.zeros(E, np.uint8)
        for e, (o, lid, rev, nn) in enumerate(outs):
            mism["obs"] += int((bits(obs[e]) != bits(o["obs"])).any())
            mism["reward"] += int((bits(rew[e]) != bits(o["reward"])).any())
            mism["status"] += int((st[e] != o["status"]).any())
            mism["flags"] += int(bool(term[e]) != o["terminated"] or bool(trunc[e]) != o["truncated"])
            for a in range(N):
                d = lid[a]
                k = np.where(d >= 250.0, 0, d / 4.0).astype(np.int64)
                mism["lidar"] += int((hit[e, a, : len(d)].astype(np.int64) != k).any())
            if rev is not None:
                got = (int(ev[e][0]), int(ev[e][1]), int(ev[e][2]), int(ev[e][3]) & 0xFFFFFFFF, int(ev[e][5]))
                want = (int(rev["rng_draws"]), int(rev["spawn_route"]), int(rev["spawned"]), int(rev["removed_mask"]), int(rev["npc_count"]))
                mism["events"] += int(got != want)
                mism["npc"] += int(int(nc[e]) != nn)
            for s in o["status"]:
                hist[s] += 1
            if o["terminated"] or o["truncated"]:
                need[e] = 1
        if need.any():
            b.reset(torch.from_numpy(need).cuda())
            for e in np.nonzero(need)[0]:
                refs[e].reset()
    out = dict(config=name, checker=checker_class().__name__, envs=E, agents=N, steps=steps, agent_steps=E * N * steps,
               mismatching_env_steps=mism, status_hist=dict(zip(po.STATUS_NAMES, hist.tolist())), seconds=round(time.time() - t0, 1))
    print(json.dumps(out), flush=True)
    b.close()
    return sum(mism.values())


if __name__ == "__main__":
    names = sys.argv[1:] or list(CONFIGS)
    bad = sum(run(n) for n in names)
    sys.exit(1 if bad else 0)
