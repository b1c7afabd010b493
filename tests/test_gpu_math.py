"""Device build of the restated libm (csrc/isx_math.cuh) against the host libm the reference binds, on the GPU."""
import ctypes as C

import numpy as np
import pytest

import pyoracle as po

pytestmark = pytest.mark.gpu


def _probe(a, b):
    from marl_traffic_intersection_b200 import _lib
    lib = _lib.load_library()
    n = a.size
    outs = [np.empty(n, np.float32) for _ in range(6)]
    _lib.check(lib, lib.isx_math_probe(0, n, a.ctypes.data, b.ctypes.data, *[o.ctypes.data for o in outs]))
    return outs


def _same(x, y):
    xb, yb = x.view(np.uint32), y.view(np.uint32)
    return (xb == yb) | (np.isnan(x) & np.isnan(y))


@pytest.mark.parametrize("seed", [0, 1])
def test_device_libm_bit_exact(seed):
    rng = np.random.default_rng(seed)
    n = 1 << 20
    a = np.concatenate([
        rng.uniform(-2 * np.pi, 2 * np.pi, n // 2), rng.uniform(-0.8, 0.8, n // 4), rng.uniform(-900, 900, n // 8),
        rng.standard_normal(n // 8) * 1e4,
    ]).astype(np.float32)
    b = np.concatenate([rng.uniform(-900, 900, n // 2), rng.uniform(-1, 1, n // 4), rng.standard_normal(n // 4) * 10]).astype(np.float32)
    a[:8] = [0.0, -0.0, np.pi, -np.pi, np.float32(np.pi), 1e-30, 0.75, 120.0]
    b[:8] = [-2.7, -2.7, 0.0, -0.0, 1.0, 1e30, -1e-30, 0.0]
    u = po.oracle_unit() if po.have_oracle() else po.ref_unit()
    sn, cs, tn, at, hy, wr = _probe(a, b)
    s2, c2 = u.sincosf(a)
    assert _same(sn, s2).all() and _same(cs, c2).all()
    assert _same(tn, u.tanf(a)).all()
    assert _same(at, u.atan2f(a, b)).all()
    assert _same(hy, u.hypotf(a, b)).all()
    pi = np.float32(3.14159265358979323846)
    w = u.fmodf(a + pi, np.full_like(a, 2 * pi))
    w = np.where(w < 0, w + np.float32(2) * pi, w) - pi
    assert _same(wr, w.astype(np.float32)).all()
