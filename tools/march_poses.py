"""Ego poses of 24 C5 envs along their episodes (CPU oracle rollout with the Philox action stream) -> gpurun_out/poses.npy"""
import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import pyoracle as po
R3 = [("IN_1", "OUT_4"), ("IN_2", "OUT_8"), ("IN_3", "OUT_12"), ("IN_4", "OUT_7"), ("IN_5", "OUT_11"), ("IN_6", "OUT_3"), ("IN_7", "OUT_10"), ("IN_8", "OUT_2")]
poses, others = [], []
for e in range(24):
    env = po.OracleEnv(num_lanes=3, ego_routes=R3, traffic=True, density=1.0, lidar_rays=72, seed=0, env_id=e)
    env.rollout(300 + 37 * e)          # the on-host Philox action stream, as the bench uses
    for s in range(40):
        env.rollout(13)
        eg = env.egos(); nn = env.npcs()
        cars = [(c["x"], c["y"], c["heading"]) for c in eg if c["alive"]] + [(c["x"], c["y"], c["heading"]) for c in nn]
        for i, c in enumerate(eg):
            if c["alive"]:
                poses.append((c["x"], c["y"], c["heading"]))
                others.append(cars)
np.save("gpurun_out/poses.npy", np.array(poses, np.float32))
print(len(poses))
