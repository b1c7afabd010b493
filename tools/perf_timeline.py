#!/usr/bin/env python
"""Per-kernel device time along one 2000-step episode of the bench workload (CUDA events around every launch).
   python tools/perf_timeline.py [envs] [buckets] [steps_per_bucket]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import CONFIGS  # noqa: E402
from marl_traffic_intersection_b200 import BatchedIntersectionEnv  # noqa: E402

N_AGENTS, ROUTES8 = CONFIGS["C5"]["agents"], CONFIGS["C5"]["routes"]
E = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
B = int(sys.argv[2]) if len(sys.argv) > 2 else 20
S = int(sys.argv[3]) if len(sys.argv) > 3 else 100
env = BatchedIntersectionEnv({"num_envs": E, "num_agents": N_AGENTS, "num_lanes": 3, "ego_routes": ROUTES8, "traffic_flow": True,
                              "traffic_density": 1.0, "lidar_rays": 72, "max_steps": 2000, "auto_reset": True, "seed": 0})
WARM = int(os.environ.get('WARM', '0'))
env.rollout(3)
torch.cuda.synchronize()
env.reset()
if WARM:
    env.rollout(WARM)
    torch.cuda.synchronize()
tot = [0.0] * 4
for b in range(B):
    ms = env.rollout_timed4(S)
    tot = [a + x for a, x in zip(tot, ms)]
    npc = env.buf["npc_count"].float().mean().item()
    us = [1e3 * x / S for x in ms]
    print(f"steps {b * S:5d}-{(b + 1) * S:5d}: traffic {us[0]:6.1f}  ego {us[1]:6.1f}  features {us[2]:6.1f}  rays {us[3]:6.1f}  "
          f"sum {sum(us):7.1f} us  -> {E * N_AGENTS / (sum(us) * 1e-6):.3e} agent-steps/s   mean NPCs {npc:.2f}")
n = B * S
us = [1e3 * x / n for x in tot]
print(f"episode mean: traffic {us[0]:.1f}  ego {us[1]:.1f}  features {us[2]:.1f}  rays {us[3]:.1f} us  -> "
      f"{E * N_AGENTS / (sum(us) * 1e-6):.3e} agent-steps/s")
print(env.stats())
