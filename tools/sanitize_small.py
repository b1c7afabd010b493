"""Smallest end-to-end exercise of every kernel for compute-sanitizer (memcheck): a few envs, traffic on, both lidar
modes, masked reset, snapshot/restore, host step, stats."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from marl_traffic_intersection_b200 import BatchedIntersectionEnv
routes = [("IN_6", "OUT_2"), ("IN_4", "OUT_8"), ("IN_5", "OUT_7")]
for rays, E in ((72, 5), (96, 1030)):
    b = BatchedIntersectionEnv(dict(num_envs=E, num_agents=3, ego_routes=routes, use_team_reward=True, traffic_flow=True, traffic_density=20.0,
                                    lidar_rays=rays, seed=1, auto_reset=True, max_steps=25))
    b.rollout(40)
    snap = b.snapshot()
    a = torch.rand(E, 3, 2, device="cuda") * 2 - 1
    for _ in range(5):
        b.step(a)
    m = torch.zeros(E, dtype=torch.uint8, device="cuda"); m[::2] = 1
    b.reset(m); b.restore(snap, m); b.restore(snap); b.observe()
    b.step_host(np.zeros((E, 3, 2), np.float32))
    print(rays, E, b.stats()["agent_steps"])
    b.close()
print("sanitize run ok")
