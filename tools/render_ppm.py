#!/usr/bin/env python
"""Write the headless picture of one env (isx_render) to a binary PPM after a short random rollout.

  python tools/render_ppm.py out.ppm [--steps 200] [--agents 6] [--density 2.0] [--lanes 3]
"""
import argparse
import sys

sys.path.insert(0, ".")
from marl_traffic_intersection_b200 import BatchedIntersectionEnv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--agents", type=int, default=6)
    ap.add_argument("--density", type=float, default=2.0)
    ap.add_argument("--lanes", type=int, default=3)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    env = BatchedIntersectionEnv({"num_envs": 1, "num_agents": a.agents, "num_lanes": a.lanes, "traffic_flow": a.density > 0,
                                  "traffic_density": a.density, "seed": a.seed, "auto_reset": True})
    env.rollout(a.steps)
    img = env.render(0).cpu().numpy()
    with open(a.out, "wb") as f:
        f.write(b"P6\n750 750\n255\n")
        f.write(img.tobytes())
    print(f"wrote {a.out}: {env.stats()}")


if __name__ == "__main__":
    main()
