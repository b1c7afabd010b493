#!/usr/bin/env python
"""Latency of ONE step through the single-env, reference-compatible facade (env.py surface) — the small-batch floor."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from marl_traffic_intersection_b200 import IntersectionEnv  # noqa: E402

for cfg in ({"num_agents": 3, "use_team_reward": True, "ego_routes": [("IN_6", "OUT_2"), ("IN_4", "OUT_8"), ("IN_5", "OUT_7")]},
            {"traffic_flow": True, "traffic_density": 0.5}):
    env = IntersectionEnv(cfg)
    env.reset()
    a = np.zeros((env.num_agents, 2), np.float32)
    a[:, 0] = 0.3
    for _ in range(50):
        env.step(a)
    t0 = time.perf_counter()
    n = 500
    for _ in range(n):
        o, r, te, tr, info = env.step(a)
        if te or tr:
            env.reset()
    print(f"{cfg}: {(time.perf_counter() - t0) / n * 1e6:.1f} us per facade step")
    env.close()
