#!/usr/bin/env python
"""Do DMA writes from the GPU and the host expander's non-temporal stores share one ceiling?  Runs the obs-row expander on T
threads and a pinned D2H copy loop alone and together, and prints the bandwidth of each (host-buffer step design aid)."""
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from marl_traffic_intersection_b200 import _lib  # noqa: E402

lib = _lib.load_library()
n, R = 65536 * 8, 72
rec = np.random.rand(n, 32).astype(np.float32)
rec[:, 31] = 1
hits = np.random.randint(0, 63, (n, R)).astype(np.uint8)
buf = np.zeros(n * 127 + 8, np.float32)
off = (-(buf.ctypes.data) // 4) % 8
dst = buf[off:off + n * 127]
T = int(sys.argv[1]) if len(sys.argv) > 1 else 12
cuts = [(n * i // T) & ~7 for i in range(T)] + [n]
stop = False
rows_done = [0] * T


def work(i):
    a, b = cuts[i], cuts[i + 1]
    while not stop:
        lib.isx_expand_obs_rows(rec[a:].ctypes.data, hits[a:].ctypes.data, R, dst[a * 127:].ctypes.data, b - a)
        rows_done[i] += b - a


nbytes = 266 * 1000 * 1000
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()


def d2h(seconds):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    k = 0
    while time.perf_counter() - t0 < seconds:
        host.copy_(dev, non_blocking=True)
        torch.cuda.synchronize()
        k += 1
    return k * nbytes / (time.perf_counter() - t0) / 1e9


print(f"D2H alone: {d2h(2.0):.1f} GB/s")
th = [threading.Thread(target=work, args=(i,)) for i in range(T)]
for t in th:
    t.start()
time.sleep(0.5)
r0, t0 = sum(rows_done), time.perf_counter()
time.sleep(2.0)
r1, t1 = sum(rows_done), time.perf_counter()
print(f"expander alone (T={T}): {(r1 - r0) * 508 / (t1 - t0) / 1e9:.1f} GB/s written")
r0, t0 = sum(rows_done), time.perf_counter()
bw = d2h(2.0)
r1, t1 = sum(rows_done), time.perf_counter()
print(f"together: D2H {bw:.1f} GB/s + expander {(r1 - r0) * 508 / (t1 - t0) / 1e9:.1f} GB/s written")
stop = True
for t in th:
    t.join()
