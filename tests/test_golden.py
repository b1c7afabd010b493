"""The C restatement (oracle/isx_oracle.c) against the committed golden fixtures, which were produced by the
reference's own C++ (tests/golden/make_golden.py).  CPU only; this is what pins the oracle where /root/reference and
oracle/_ref are absent."""
import os
import sys

import numpy as np
import pytest

import pyoracle as po

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden as mg  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
pytestmark = pytest.mark.skipif(not po.have_oracle(), reason="oracle/libisx_oracle.so not built (run __graft_entry__.build())")


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype == np.float32:
        return (a.view(np.uint32) == b.view(np.uint32)).all()
    return (a == b).all()


@pytest.mark.parametrize("name", list(mg.CASES))
def test_oracle_reproduces_reference_rollout(name):
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    got = mg.run(po.OracleEnv, mg.CASES[name])
    for k in gold.files:
        if k == "events":   # the reference driver cannot see collided_mask; the compared columns are all inferable
            assert _same(gold[k], got[k]), k
        else:
            assert _same(gold[k], got[k]), f"{name}: {k} differs from the reference fixture"


def test_oracle_geometry_and_routes_match_reference():
    gold = np.load(os.path.join(GOLD, "geometry_routes.npz"))
    u = po.oracle_unit()
    for L in (2, 3):
        assert (np.packbits(u.road_map(L)) == gold[f"road_{L}"]).all()
        assert (np.packbits(u.line_map(L)) == gold[f"line_{L}"]).all()
        ids = [f"IN_{k}" for k in range(1, 4 * L + 1)] + [f"OUT_{k}" for k in range(1, 4 * L + 1)]
        i = 0
        for a in ids:
            for b in ids:
                n, p, intent, sp = u.route(L, a, b)
                assert n == 160 and _same(p, gold[f"paths_{L}"][i])
                assert [intent, *sp.view(np.uint32).tolist()] == gold[f"meta_{L}"][i].tolist()
                i += 1


def test_survey_known_answers():
    """SURVEY.md §4 KATs, observed on the reference during the survey."""
    u = po.oracle_unit()
    n, path, intent, sp = u.route(3, "IN_6", "OUT_2")
    assert (n, intent) == (160, 2) and sp[0] == 720.0 and sp[1] == 270.0 and sp.view(np.uint32)[2] == 0xC0490FDB
    assert tuple(path[50]) == (585.0, 270.0) and abs(path[159][0] - 438.84) < 1e-3 and abs(path[159][1] - 32.7) < 1e-3
    e = po.OracleEnv(3, [("IN_6", "OUT_2")])
    crash_steps = []
    for s in range(1, 400):
        r = e.step([[0.5, 0.0]])
        c = e.egos()[0]
        if s == 1:
            assert c["x"] == np.float32(719.875) and c["v"] == np.float32(0.125)
            assert r["reward"].view(np.uint32)[0] == 0xBC75C28F
        if s == 60:
            assert c["x"] == np.float32(491.25) and c["v"] == np.float32(7.5)
            assert e.lidar(0)[2:6].tolist() == [160.0, 108.0, 88.0, 76.0]
        if r["status"][0] == po.STATUS_NAMES.index("CRASH_WALL"):
            crash_steps.append(s)
            assert abs(r["reward"][0] + 5.068404) < 1e-5
    assert crash_steps == [131, 262, 393]
    d = e.lidar(0)
    assert set(np.unique(d)).issubset(set(np.arange(4, 252, 4).astype(np.float32)) | {np.float32(250.0)})


def test_numpy_rollout_histogram_kat():
    """config 1 / config 2, numpy default_rng(0) actions, 2000 steps."""
    for routes, team, expect in (
        ([("IN_6", "OUT_2")], False, {"ALIVE": 1995, "CRASH_LINE": 3, "CRASH_WALL": 2}),
        ([("IN_6", "OUT_2"), ("IN_4", "OUT_8"), ("IN_5", "OUT_7")], True, {"ALIVE": 5981, "CRASH_CAR": 6, "CRASH_WALL": 8, "CRASH_LINE": 5}),
    ):
        e = po.OracleEnv(3, routes, use_team=team)
        rng = np.random.default_rng(0)
        cnt = {}
        for _ in range(2000):
            r = e.step(rng.uniform(-1, 1, (len(routes), 2)).astype(np.float32))
            for s in r["status"]:
                cnt[po.STATUS_NAMES[s]] = cnt.get(po.STATUS_NAMES[s], 0) + 1
        assert cnt == expect


def test_libm_points_match_this_machine():
    """The libm values recorded with the fixtures equal this machine's libm (same glibc algorithms); if this fails the
    host's libm differs from the one the fixtures were made with and float-bit fixtures cannot be expected to match."""
    g = np.load(os.path.join(GOLD, "libm_points.npz"))
    u = po.oracle_unit()
    s, c = u.sincosf(g["a"])
    assert _same(s, g["sin"]) and _same(c, g["cos"]) and _same(u.tanf(g["a"]), g["tan"])
    assert _same(u.atan2f(g["a"], g["b"]), g["atan2"]) and _same(u.hypotf(g["a"], g["b"]), g["hypot"])


def test_neighbor_sort_restatements_match_recorded_std_sort():
    """The oracle's (and, when built, the product's host build of the) libstdc++ std::sort restatement against rank
    orders recorded from the real std::sort — ties, > 16 elements, and adversary inputs that hit the heap fallback."""
    g = np.load(os.path.join(GOLD, "sort_vectors.npz"))
    units = [po.oracle_unit()]
    hu = os.path.join(os.path.dirname(GOLD), "..", "marl-traffic-intersection_b200", "csrc", "libisx_host_units.so")
    if os.path.exists(hu):
        units.append(po.unit_of(hu, "isxh_"))
    off = 0
    fell_back = 0
    for n in g["lens"]:
        k, want = g["keys"][off:off + n], g["perms"][off:off + n]
        off += n
        for u in units:
            got, heaps = u.std_sort(k)
            assert (got == want).all(), (int(n), k.tolist())
            fell_back += int(heaps or 0) > 0
    assert fell_back > 0
