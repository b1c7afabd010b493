#!/usr/bin/env python
"""Per-kernel device time along one 2000-step episode of the bench workload (CUDA events around every launch).
   python tools/perf_timeline.py [envs] [buckets] [steps_per_bucket]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import N_AGENTS, ROUTES8  # noqa: E402
from marl_traffic_intersection_b200 import BatchedIntersectionEnv  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
B = int(sys.argv[2]) if len(sys.argv) > 2 else 20
S = int(sys.argv[3]) if len(sys.argv) > 3 else 100
env = BatchedIntersectionEnv({"num_envs": E, "num_agents": N_AGENTS, "num_lanes": 3, "ego_routes": ROUTES8, "traffic_flow": True,
                              "traffic_density": 1.0, "lidar_rays": 72, "max_steps": 2000, "auto_reset": True, "seed": 0})
env.rollout(3)
torch.cuda.synchronize()
env.reset()
tot_d = tot_l = 0.0
for b in range(B):
    d, l = env.rollout_timed(S)
    tot_d += d
    tot_l += l
    npc = env.buf["npc_count"].float().mean().item()
    print(f"steps {b * S:5d}-{(b + 1) * S:5d}: k_dynamics {1e3 * d / S:7.1f} us  k_lidar_obs {1e3 * l / S:7.1f} us  "
          f"sum {1e3 * (d + l) / S:7.1f} us  -> {E * N_AGENTS / ((d + l) / S * 1e-3):.3e} agent-steps/s   mean NPCs {npc:.2f}")
n = B * S
print(f"episode mean: k_dynamics {1e3 * tot_d / n:.1f} us  k_lidar_obs {1e3 * tot_l / n:.1f} us  -> "
      f"{E * N_AGENTS / ((tot_d + tot_l) / n * 1e-3):.3e} agent-steps/s")
print(env.stats())
