"""How many exact sample tests a beam needs after the analytic jump, on ego poses of a C5 rollout (tools/march_poses.py writes
gpurun_out/poses.npy): python tools/march_poses.py && python tools/march_stats.py"""
import ctypes as C, numpy as np, sys
lib = C.CDLL("/root/repo/marl-traffic-intersection_b200/csrc/libisx_host_units.so")
P = np.load("gpurun_out/poses.npy")
R = 72
deg = -180.0 + np.arange(R, dtype=np.float32) * np.float32(360.0 / (R - 1))
rel = (deg.astype(np.float32) * np.float32(np.pi) / np.float32(180.0)).astype(np.float32)
cx = np.repeat(P[:, 0], R); cy = np.repeat(P[:, 1], R)
ang = (P[:, 2:3] + rel[None, :]).astype(np.float32).ravel()
n = len(cx)
out = np.zeros((n, 6), np.int32)
f = lambda a: np.ascontiguousarray(a, np.float32)
cx, cy, ang = f(cx), f(cy), f(ang)
lib.isxh_road_events(3, n, cx.ctypes.data_as(C.c_void_p), cy.ctypes.data_as(C.c_void_p), ang.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
bad = ((out[:, 0] != out[:, 2]) | (out[:, 1] != out[:, 3])) & ~((out[:, 1] == 0) & (out[:, 3] == 0))
print("rays", n, "mismatch", bad.sum())
t = out[:, 4]
print("tests: mean %.2f" % t.mean(), "hist:", np.bincount(np.minimum(t, 12))[:13] / n)
for k in range(0, 8): print(f" open after {k} tests: {(t > k).mean()*100:.2f}%")
# per 32-beam piece: max tests and number of open rays after 3 tests
pieces = t[: n // 32 * 32].reshape(-1, 32)
for k in (2, 3, 4):
    op = (pieces > k).sum(1)
    print(f"after {k} lockstep tests: pieces with open rays {(op > 0).mean()*100:.1f}%, mean open/piece {op.mean():.2f}; remaining samples of open rays p50 {np.median(pieces[pieces>k]-k):.0f} mean {(pieces[pieces>k]-k).mean():.1f}")
print("hit fraction", out[:, 3].mean(), "ksafe==62 fraction", (out[:, 5] >= 62).mean())
