#!/usr/bin/env python
"""Run-to-run determinism stress: N pairs of identical batches, different call granularity; prints the buffers that differ."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
import pyoracle as po  # noqa: E402
from marl_traffic_intersection_b200 import BatchedIntersectionEnv  # noqa: E402

R3 = po.ROUTES_3LANES
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
cfg = dict(num_envs=4096, num_agents=8, num_lanes=3, ego_routes=R3[:8], traffic_flow=True, traffic_density=2.0, lidar_rays=72,
           seed=77, auto_reset=True, max_steps=120)
bad = 0
for r in range(reps):
    a, b = BatchedIntersectionEnv(cfg), BatchedIntersectionEnv(cfg)
    a.rollout(300)
    for _ in range(3 + r):
        b.rollout(300 // (3 + r))
    b.rollout(300 - (300 // (3 + r)) * (3 + r))
    torch.cuda.synchronize()
    for k in a.buf:
        x, y = a.buf[k].cpu().numpy(), b.buf[k].cpu().numpy()
        if k.startswith("npc_") and k != "npc_count":
            n = a.buf["npc_count"].cpu().numpy()
            m = np.arange(x.shape[1])[None, :] < n[:, None]
            x, y = np.where(m, x, 0), np.where(m, y, 0)
        d = (x.view(np.uint8) != y.view(np.uint8))
        if d.any():
            bad += 1
            idx = np.argwhere(x != y)
            print(f"rep {r}: {k} differs in {int((x != y).sum())} elements, first at {idx[0].tolist()}: {x[tuple(idx[0])]} vs {y[tuple(idx[0])]}")
    a.close(); b.close()
print("mismatching buffers:", bad)
