"""HOST build of the product's entity-level arithmetic (csrc/isx_sim.cuh, isx_tables.h — the same source the kernels
compile) against the CPU checker, on the CPU.  Bit-exact everywhere."""
import os

import numpy as np
import pytest

import pyoracle as po

HU = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "marl-traffic-intersection_b200", "csrc", "libisx_host_units.so")
pytestmark = pytest.mark.skipif(not (os.path.exists(HU) and (po.have_ref() or po.have_oracle())), reason="host-unit library or checker not built")


def checker():
    return po.ref_unit() if po.have_ref() else po.oracle_unit()


def host():
    return po.unit_of(HU, "isxh_")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("lanes", [2, 3])
def test_folded_bitmap_line_mask_and_routes(lanes):
    c, h = checker(), host()
    assert (c.road_map(lanes) == h.road_map(lanes)).all()       # exhaustive 750x750 vs is_on_road
    assert (c.line_map(lanes) == h.line_map(lanes)).all()       # exhaustive 750x750 vs LineMask
    ids = [f"IN_{k}" for k in range(1, 4 * lanes + 1)] + [f"OUT_{k}" for k in range(1, 4 * lanes + 1)]
    for a in ids:
        for b in ids:
            n1, p1, i1, s1 = c.route(lanes, a, b)
            n2, p2, i2, s2 = h.route(lanes, a, b)
            assert n1 == n2 and i1 == i2 and (bits(p1) == bits(p2)).all() and (bits(s1) == bits(s2)).all()
    assert h.route(lanes, "IN_99", "OUT_1")[0] == -1 and h.route(lanes, "IN_1", "OUT_99")[0] == -2
    assert h.route(lanes, "IN_01", "OUT_1")[0] == -1 and h.route(lanes, "in_1", "OUT_1")[0] == -1


def test_float_geometry_probes():
    c, h = checker(), host()
    rng = np.random.default_rng(0)
    for _ in range(20000):
        x, y = rng.uniform(-150, 900, 2)
        if rng.random() < 0.3:
            x = 375 + rng.choice([-1, 1]) * (126 + rng.uniform(-0.01, 0.01))
        L = int(rng.choice([2, 3]))
        assert c.on_road(L, x, y) == h.on_road(L, x, y)
        assert c.yellow(L, x, y) == h.yellow(L, x, y)


def test_car_update_and_sat():
    c, h = checker(), host()
    rng = np.random.default_rng(1)
    for _ in range(30000):
        s = np.array([rng.uniform(-50, 800), rng.uniform(-50, 800), rng.uniform(0, 8), rng.uniform(-3.2, 3.2), 0, rng.uniform(-0.7, 0.7)], np.float32)
        thr = np.float32(rng.choice([0.0, rng.uniform(-1, 1)]))
        st = np.float32(rng.uniform(-2.5, 2.5))
        assert (bits(c.car_update(s, thr, st, 1 / 60)) == bits(h.car_update(s, thr, st, 1 / 60))).all()
    hits = 0
    for _ in range(30000):
        a = np.array([rng.uniform(300, 400), rng.uniform(300, 400), rng.uniform(-3.2, 3.2)], np.float32)
        d, th = rng.uniform(0, 90), rng.uniform(0, 6.3)
        b = np.array([a[0] + d * np.cos(th), a[1] + d * np.sin(th), rng.uniform(-3.2, 3.2)], np.float32)
        r = c.collide(a, b)
        hits += r
        assert r == h.collide(a, b)
        assert (bits(c.corners(a)) == bits(h.corners(a))).all()
    assert 5000 < hits < 25000


def test_accelerated_lidar_equals_reference_march():
    """Skip-table + strip-box + slab/verify lidar == the reference's sample-by-sample march, incl. off-screen egos,
    integer poses, a copy of self in the list, both beam counts."""
    c, h = checker(), host()
    rng = np.random.default_rng(2)
    road = {L: c.road_map(L) for L in (2, 3)}
    for it in range(6000):
        L = 3 if it % 4 else 2
        mode = it % 5
        if mode == 0:
            x, y = rng.uniform(-130, 880, 2)
        elif mode == 1:
            while True:
                x, y = rng.uniform(0, 750, 2)
                if road[L][int(y), int(x)]:
                    break
        elif mode == 2:
            x, y = rng.integers(0, 750, 2).astype(float)
        elif mode == 3:
            x, y = 375 + rng.uniform(-130, 130), rng.uniform(0, 750)
        else:
            x, y = rng.uniform(0, 750), 375 + rng.uniform(-130, 130)
        hd = rng.uniform(-np.pi, np.pi) if it % 7 else rng.choice([0, np.pi / 2, -np.pi / 2, np.pi, -np.pi])
        others = []
        for j in range(rng.integers(0, 12)):
            d = rng.uniform(0, 300) if j % 3 else rng.uniform(20, 80)
            th = rng.uniform(0, 6.3)
            others.append([x + d * np.cos(th), y + d * np.sin(th), rng.uniform(-np.pi, np.pi) if j % 2 else rng.choice([0, np.pi / 2, np.pi, -np.pi / 2])])
        if it % 11 == 0 and others:
            others[0] = [x, y, hd]
        rays = 96 if it % 3 else 72
        a, b = c.lidar(L, rays, [x, y, hd], others), h.lidar(L, rays, [x, y, hd], others)
        assert (a == b).all(), (L, rays, x, y, hd, others)


def _tie_heavy_keys(rng, n):
    kind = rng.integers(0, 5)
    if kind == 0:
        return rng.integers(0, 3, n).astype(np.float32)                   # two or three distinct values
    if kind == 1:
        return rng.integers(0, max(n // 2, 1), n).astype(np.float32)      # many pairs
    if kind == 2:
        k = rng.random(n).astype(np.float32)
        k[rng.integers(0, n, n // 3 + 1)] = k[0]                          # one value repeated among distinct ones
        return k
    if kind == 3:
        return np.sort(rng.integers(0, 8, n)).astype(np.float32)[:: (1 if rng.random() < 0.5 else -1)].copy()
    return rng.random(n).astype(np.float32)                               # no ties


def test_neighbor_sort_matches_the_toolchain_std_sort():
    """IntersectionEnv.cpp:490 is an UNSTABLE std::sort: with equal distances and > 16 neighbours the outcome is the
    library's introsort.  Product restatement (isx_sim.cuh stdsort) == oracle restatement == real std::sort."""
    if not po.have_ref():
        pytest.skip("needs the compiled reference for std::sort itself")
    r, o, h = po.ref_unit(), po.oracle_unit(), host()
    rng = np.random.default_rng(7)
    for n in list(range(0, 70)) * 30:
        keys = _tie_heavy_keys(rng, n) if n else np.zeros(0, np.float32)
        pr, _ = r.std_sort(keys)
        ph, _ = h.std_sort(keys)
        pq, _ = o.std_sort(keys)
        assert (pr == ph).all() and (pr == pq).all(), (n, keys.tolist(), pr.tolist(), ph.tolist())
        assert sorted(pr.tolist()) == list(range(n)) and (np.diff(keys[pr]) >= 0).all()


def test_neighbor_sort_heap_fallback_against_adversary():
    """McIlroy's adversary drives std::sort to its depth limit; the heap-sort fallback must then agree too."""
    if not po.have_ref():
        pytest.skip("needs the compiled reference for std::sort itself")
    r, o, h = po.ref_unit(), po.oracle_unit(), host()
    fell_back = 0
    for n in range(17, 64):
        base = r.sort_adversary(n)
        for q in (1, 2, 3):                                               # q > 1 folds the killer sequence into ties
            keys = np.floor(base / q).astype(np.float32)
            pr, _ = r.std_sort(keys)
            ph, nh = h.std_sort(keys)
            pq, nq = o.std_sort(keys)
            assert (pr == ph).all() and (pr == pq).all(), (n, q)
            assert nh == nq
            fell_back += nh > 0
    assert fell_back > 0


def test_host_obs_row_expander_matches_the_device_formula():
    """isx_expand_obs_rows (the host half of the compact obs transport of isx_step_host): rows rebuilt from 32-float records
    + u8 hit indices must equal what k_lidar_obs writes: obs[31+i] = float(4k) * (1/250), k = 0 -> 250 * (1/250); columns
    behind the last beam 0; dead ego -> all-zero row.  Every alignment of the destination, every tail length."""
    import ctypes as C
    from marl_traffic_intersection_b200 import _lib
    lib = _lib.load_library()
    rng = np.random.default_rng(0)
    inv = np.float32(1.0) / np.float32(250.0)
    for R in (72, 96, 33, 1):
        for n in (0, 1, 7, 8, 9, 64, 1001):
            rec = rng.normal(size=(n, 32)).astype(np.float32)
            alive = rng.random(n) > 0.2
            rec[:, 31] = alive
            hits = rng.integers(0, 63, size=(n, R)).astype(np.uint8)
            hits[rng.random((n, R)) < 0.4] = 0
            want = np.zeros((n, 127), np.float32)
            want[:, :31] = rec[:, :31]
            lid = (4 * hits.astype(np.int32)).astype(np.float32) * inv
            lid[hits == 0] = np.float32(250.0) * inv
            want[:, 31:31 + R] = lid
            want[~alive] = 0.0
            for off in (0, 1, 3, 5):                          # destination misaligned by off floats
                buf = np.full(n * 127 + 16, np.float32(np.nan))
                dst = buf[off:off + n * 127]
                assert lib.isx_expand_obs_rows(rec.ctypes.data, hits.ctypes.data, R, dst.ctypes.data, n) == 0
                assert (dst.view(np.uint32) == want.reshape(-1).view(np.uint32)).all(), (R, n, off)
                assert np.isnan(buf[:off]).all() and np.isnan(buf[off + n * 127:]).all()


def _road_events(lanes, cx, cy, ang):
    import ctypes as C
    lib = C.CDLL(HU)
    n = len(cx)
    out = np.zeros((n, 6), np.int32)
    f = [np.ascontiguousarray(a, np.float32) for a in (cx, cy, ang)]
    lib.isxh_road_events(lanes, n, *[a.ctypes.data_as(C.c_void_p) for a in f], out.ctypes.data_as(C.c_void_p))
    return out


@pytest.mark.parametrize("lanes", [1, 2, 3, 4])
def test_analytic_road_bound_keeps_the_march_exact(lanes):
    """k_lidar_obs jumps over the samples ray_safe_samples proves free of road events (rounded-quadrant entry time) and
    tests the following samples exactly.  Same first event as the skip-table march (itself pinned to Lidar.cpp's
    sample-by-sample loop above) on uniform, axis-aligned, wall-hugging, arc-tangent and grazing rays; a no-hit ray may
    report a later break index (cars are only tested against on-screen pixels), never a different hit."""
    import ctypes as C
    assert C.CDLL(HU).isxh_ana_enabled(lanes) == 1
    rng = np.random.default_rng(100 + lanes)
    n = 150_000
    rw, U = 42 * lanes, 42 * lanes + 84
    side = rng.choice([-1, 1], n)
    eps = rng.choice([0, .01, -.01, .5, -.5, 1, -1, 1.4, -1.4, 1.6, -1.6, 3, -3], n)
    th = rng.uniform(0, 2 * np.pi, n)
    sx, sy = rng.choice([-1, 1], n), rng.choice([-1, 1], n)
    rr = 84 + eps + rng.uniform(-.2, .2, n)
    ox, oy = 375 + sx * U + rr * np.cos(th), 375 + sy * U + rr * np.sin(th)
    px, py = rng.uniform(200, 550, n), rng.uniform(200, 550, n)
    tx, ty = 375 + sx * U + (84 + rng.normal(0, 1, n)) * np.cos(th), 375 + sy * U + (84 + rng.normal(0, 1, n)) * np.sin(th)
    ax = rng.choice([0, np.pi / 2, np.pi, -np.pi / 2, -np.pi, 2 * np.pi], n) + rng.choice([0, 1e-7, -1e-7, 1e-4, -1e-4, 1e-2, -1e-2], n)
    sets = {
        "uniform": (rng.uniform(-30, 780, n), rng.uniform(-30, 780, n), rng.uniform(-7, 7, n)),
        "axis": (rng.integers(0, 750, n) + rng.choice([0, .5, .999, .001], n), rng.integers(0, 750, n) + rng.choice([0, .5, .999, .001], n), ax),
        "wall": (375 + side * (rw + eps), rng.uniform(0, 750, n), rng.uniform(-np.pi, np.pi, n)),
        "arc": (ox, oy, rng.uniform(-np.pi, np.pi, n)),
        "arc_tangent": (ox, oy, -(th + np.pi / 2 * rng.choice([-1, 1], n)) + rng.normal(0, .02, n)),
        "aim_arc": (px, py, np.arctan2(-(ty - py), tx - px)),
        "graze": (rng.uniform(0, 750, n), 375 + side * (rw - rng.uniform(0, 8, n)), rng.choice([0, np.pi], n) + rng.normal(0, .03, n)),
    }
    for name, (x, y, a) in sets.items():
        o = _road_events(lanes, x, y, a)
        same_hit = (o[:, 1] == o[:, 3]) & ((o[:, 0] == o[:, 2]) | (o[:, 1] == 0))
        assert same_hit.all(), (name, lanes, np.nonzero(~same_hit)[0][:3])
        assert (o[:, 5] <= np.maximum(o[:, 0] - 1, 0))[o[:, 1] == 1].all(), name      # the jump never passes a hit


def test_path_index_near_far_window_is_exact():
    """Car::update_path_index with the far part of the 50-point window skipped under the triangle-inequality bound
    (path_far_table) == the plain scan, for cars on, near and far from their route."""
    import ctypes as C
    lib = C.CDLL(HU)
    rng = np.random.default_rng(7)
    for lanes in (2, 3):
        for a, b in po.default_routes(lanes):
            r = host().route(lanes, a, b)
            if r[1] is None:
                continue
            path, n = r[1], 40_000
            idx = rng.integers(-1, 160, n).astype(np.int32)
            base = path[np.clip(idx, 0, 159)]
            scale = rng.choice([0.5, 5, 20, 60, 200, 600], n)
            x = (base[:, 0] + rng.normal(0, 1, n) * scale).astype(np.float32)
            y = (base[:, 1] + rng.normal(0, 1, n) * scale).astype(np.float32)
            fast, full = np.zeros(n, np.int32), np.zeros(n, np.int32)
            rc = lib.isxh_path_index(lanes, a.encode(), b.encode(), n, *[v.ctypes.data_as(C.c_void_p) for v in (idx, x, y, fast, full)])
            assert rc == 0 and (fast == full).all(), (lanes, a, b)
