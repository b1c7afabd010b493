#!/usr/bin/env python
"""Host-side throughput of the obs-row expander (isx_expand_obs_rows) alone: T Python threads (ctypes releases the GIL), each on
its own slice of a 65,536 x 8-agent batch.  Tells how much of the host-buffer step's time the host memory system can take."""
import sys
import threading
import time

import numpy as np

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
for T in (1, 2, 4, 6, 8, 12, 16):
    cuts = [(n * i // T) & ~7 for i in range(T)] + [n]

    def work(i):
        a, b = cuts[i], cuts[i + 1]
        lib.isx_expand_obs_rows(rec[a:].ctypes.data, hits[a:].ctypes.data, R, dst[a * 127:].ctypes.data, b - a)
    best = 1e9
    for _ in range(5):
        th = [threading.Thread(target=work, args=(i,)) for i in range(T)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        best = min(best, time.perf_counter() - t0)
    print(f"T={T:2d}: {best * 1e3:7.2f} ms per 524,288 rows  -> {n * 508 / best / 1e9:6.1f} GB/s written, {n * (128 + R) / best / 1e9:5.1f} GB/s read")
