#!/usr/bin/env python
"""Do the latency-bound small kernels of one env range hide behind the issue-bound beam kernel of another?  Two handles
of E/2 envs step concurrently on two streams, with k_lidar_obs capped at ISX_LIDAR_CTAS_PER_SM CTAs per SM so that the
other stream's kernels can be co-resident; compared with one handle of E envs on one stream."""
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from bench import R3  # noqa: E402
N_AGENTS, ROUTES8 = 8, R3[:8]
from marl_traffic_intersection_b200 import BatchedIntersectionEnv  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
K = 100


def mk(n, base):
    return BatchedIntersectionEnv({"num_envs": n, "num_agents": N_AGENTS, "num_lanes": 3, "ego_routes": ROUTES8, "traffic_flow": True,
                                   "traffic_density": 1.0, "lidar_rays": 72, "max_steps": 2000, "auto_reset": True, "seed": 0, "env_id_base": base})


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    return time.perf_counter() - t0


one = mk(E, 0)
one.rollout(300)
t1 = timed(lambda: one.rollout(K))
print(f"one handle, {E} envs: {t1 / K * 1e6:.1f} us/step -> {E * N_AGENTS * K / t1:.3e} agent-steps/s  (ISX_LIDAR_CTAS_PER_SM={os.environ.get('ISX_LIDAR_CTAS_PER_SM')})")
one.close()
for parts in (2, 4):
    hs = [mk(E // parts, i * (E // parts)) for i in range(parts)]
    ss = [torch.cuda.Stream() for _ in range(parts)]
    for h, s in zip(hs, ss):
        with torch.cuda.stream(s):
            h.rollout(300)

    def run():
        for k in range(K):                      # interleave the launches so that the streams really overlap
            for h, s in zip(hs, ss):
                with torch.cuda.stream(s):
                    h.rollout(1)
    t = timed(run)
    print(f"{parts} handles x {E // parts} envs on {parts} streams: {t / K * 1e6:.1f} us/step -> {E * N_AGENTS * K / t:.3e} agent-steps/s")
    for h in hs:
        h.close()
