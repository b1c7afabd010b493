#!/usr/bin/env python
"""Where does the host-buffer step spend its time?  Prints the pinned D2H / H2D bandwidth of this box at the shard and
full-batch obs sizes, then the measured isx_step_pinned time, so the PCIe floor of the e2e metric is known."""
import sys
import time

import torch

sys.path.insert(0, ".")
from marl_traffic_intersection_b200 import BatchedIntersectionEnv  # noqa: E402


def bw(nbytes, d2h=True, reps=20):
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    a, b = (host, dev) if d2h else (dev, host)
    for _ in range(3):
        a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        a.copy_(b, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main():
    E, N = (int(sys.argv[1]) if len(sys.argv) > 1 else 8192), 8
    obs_bytes = E * N * 127 * 4
    for frac in (1, 4, 8, 16):
        print(f"D2H {obs_bytes // frac / 1e6:8.2f} MB : {bw(obs_bytes // frac):6.1f} GB/s")
    print(f"H2D {E * N * 8 / 1e6:8.2f} MB : {bw(E * N * 8, d2h=False):6.1f} GB/s")
    env = BatchedIntersectionEnv({"num_envs": E, "num_agents": N, "traffic_flow": True, "traffic_density": 1.0, "lidar_rays": 72,
                                  "seed": 0, "auto_reset": True})
    env.reset()
    env.rollout(300)
    import numpy as np
    act = np.random.default_rng(0).uniform(-1, 1, (E, N, 2)).astype(np.float32)
    for _ in range(10):
        env.step_host(act)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 100
    for _ in range(K):
        env.step_host(act)
    t = (time.perf_counter() - t0) / K
    import ctypes as C
    from marl_traffic_intersection_b200 import _lib
    t0 = time.perf_counter()
    for _ in range(K):
        _lib.check(env._lib, env._lib.isx_step_pinned(env._h, C.c_float(1.0 / 60.0), env._stream()))
    tc = (time.perf_counter() - t0) / K
    print(f"isx_step_pinned alone (no Python staging): {tc * 1e6:.1f} us/step")
    ms = (C.c_float * 64)()
    for _ in range(3):
        n = _lib.check(env._lib, env._lib.isx_pipe_timeline(env._h, C.c_float(1.0 / 60.0), env._stream(), ms, 16))
    for i in range(n):
        k0, k1, c0, c1 = (ms[4 * i + j] * 1e3 for j in range(4))
        print(f"  range {i}: kernels {k0:7.1f} -> {k1:7.1f} us   copy {c0:7.1f} -> {c1:7.1f} us")
    io = env.host_step_bytes()
    print(f"step_host: {t * 1e6:.1f} us/step  -> {E * N / t:.3e} agent-steps/s;  D2H {io['d2h'] / 1e6:.1f} MB/step, {io['host_expand_threads']} host threads, "
          f"{io['pipeline_ranges']} ranges")


if __name__ == "__main__":
    main()
