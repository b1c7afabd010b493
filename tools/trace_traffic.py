#!/usr/bin/env python
"""Per-phase latency of k_traffic for single warps (clock64 stamps; ISX_TRACE=1).  python tools/trace_traffic.py [envs]"""
import ctypes as C, os, sys
os.environ["ISX_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bench import R3
N_AGENTS, ROUTES8 = 8, R3[:8]
from marl_traffic_intersection_b200 import BatchedIntersectionEnv, _lib
E = int(sys.argv[1]) if len(sys.argv) > 1 else 512
env = BatchedIntersectionEnv({"num_envs": E, "num_agents": N_AGENTS, "num_lanes": 3, "ego_routes": ROUTES8, "traffic_flow": True,
                              "traffic_density": 1.0, "lidar_rays": 72, "max_steps": 2000, "auto_reset": True, "seed": 0})
env.rollout(400)
torch.cuda.synchronize()
tr = np.zeros((E, 16), np.int64)
_lib.check(env._lib, env._lib.isx_trace_read(env._h, tr.ctypes.data))
names = ["load+rng+spawn", "own-state phase", "sequential loop", "collisions", "erase+write"]
d = np.diff(tr[:, :6], axis=1)
c = tr[:, 6]
print("NPC count histogram:", np.bincount(c.astype(int)))
for k in sorted(set(c.tolist())):
    m = c == k
    print(f"c={k}: n={m.sum():4d}  " + "  ".join(f"{names[i]} {d[m, i].mean():8.0f}" for i in range(5)) + f"   total {d[m].sum(axis=1).mean():8.0f} cycles  (max {d[m].sum(axis=1).max()})")
m = c >= 2
q = tr[m]
print("first NPC of the loop (envs with c>=2): shfl+pair flags %.0f  ghost scan %.0f  motion update %.0f  path index %.0f cycles" % ((q[:, 8] - q[:, 2]).mean(), (q[:, 9] - q[:, 8]).mean(), (q[:, 10] - q[:, 9]).mean(), (q[:, 11] - q[:, 10]).mean()))
print("kernel span (first start -> last end):", tr[:, 5].max() - tr[:, 0].min(), "cycles")
