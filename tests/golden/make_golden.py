"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libisx_ref.so = the unmodified C++ of
/root/reference/cpp behind oracle/ref_driver.cpp).  Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

The fixtures pin the C restatement (tests/test_golden.py, CPU) and the CUDA stepper (tests/test_gpu_golden.py)
on machines where the reference sources — and possibly oracle/_ref — are absent.  Note: sincosf bits depend on the
host's glibc ifunc choice (FMA vs SSE2 build); these were produced on an FMA-capable x86-64 CPU, glibc 2.39."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import pyoracle as po  # noqa: E402

R3 = po.ROUTES_3LANES
CASES = {
    "c1_single": dict(num_lanes=3, ego_routes=[("IN_6", "OUT_2")]),
    "c2_team3": dict(num_lanes=3, ego_routes=[("IN_6", "OUT_2"), ("IN_4", "OUT_8"), ("IN_5", "OUT_7")], use_team=True),
    "c3_traffic": dict(num_lanes=3, ego_routes=[("IN_6", "OUT_2")], traffic=True, density=0.5),
    "c4_eight": dict(num_lanes=3, ego_routes=R3[:8]),
    "c5_eight_traffic72": dict(num_lanes=3, ego_routes=R3[:8], traffic=True, density=1.0, lidar_rays=72),
    "two_lanes_norespawn": dict(num_lanes=2, ego_routes=po.ROUTES_2LANES[:3], traffic=True, density=2.0, respawn=False, max_steps=250),
    # 20 egos on 12 spawn points + traffic: > 16 neighbours with exactly equal distances, the regime where
    # IntersectionEnv.cpp:490's std::sort is not stable (libstdc++ introsort decides the order)
    "stacked20_ties": dict(num_lanes=3, ego_routes=[R3[i % 12] for i in range(20)], traffic=True, density=3.0, lidar_rays=72, max_steps=150),
    "four_lanes_team": dict(num_lanes=4, ego_routes=[("IN_1", "OUT_9"), ("IN_6", "OUT_15"), ("IN_11", "OUT_2"), ("IN_16", "OUT_4")], use_team=True,
                            traffic=True, density=2.0),
}
STEPS, SEED, ENV_ID = 600, 20261018, 3


def run(env_cls, kw, steps=STEPS, seed=SEED, env_id=ENV_ID):
    e = env_cls(seed=seed, env_id=env_id, **kw)
    n = e.n
    h = hashlib.sha256()
    rew = np.zeros((steps, n), np.float32)
    status = np.zeros((steps, n), np.int8)
    flags = np.zeros((steps, 3), np.int32)
    events = np.zeros((steps, 5), np.int64)
    lidar_hash = hashlib.sha256()
    obs_keep = {}
    h.update(e.obs().tobytes())
    for t in range(steps):
        a = po.philox_actions(seed, env_id, e.tick + 1, n)
        o = e.step(a)
        h.update(o["obs"].tobytes())
        rew[t] = o["reward"]
        status[t] = o["status"]
        flags[t] = (o["terminated"], o["truncated"], o["agents_alive"])
        ev = e.events()
        events[t] = (ev["rng_draws"], ev["spawn_route"], ev["spawned"], ev["removed_mask"], ev["npc_count"])
        for i in range(n):
            lidar_hash.update(e.lidar(i).tobytes())
        if t % 100 == 99:
            obs_keep[f"obs_{t + 1}"] = o["obs"].copy()
        if o["terminated"] or o["truncated"]:
            e.reset()
    eg = e.egos()
    return dict(reward=rew, status=status, flags=flags, events=events, obs_sha256=np.frombuffer(h.digest(), np.uint8),
                lidar_sha256=np.frombuffer(lidar_hash.digest(), np.uint8),
                final_ego=np.stack([eg[f] for f in ("x", "y", "v", "heading", "steer", "prev_dist")]).astype(np.float32), **obs_keep)


def sort_vectors():
    """Key arrays and the rank order the toolchain's std::sort (libstdc++ 13, as linked into oracle/_ref) leaves them in:
    tie-heavy random keys for every n in [0, 64) and McIlroy-adversary keys that force the heap-sort fallback."""
    u = po.ref_unit()
    rng = np.random.default_rng(11)
    keys, perms = [], []
    for n in list(range(0, 64)) * 4:
        kind = rng.integers(0, 3)
        k = (rng.integers(0, 3, n) if kind == 0 else rng.integers(0, max(n // 2, 1), n) if kind == 1 else rng.random(n)).astype(np.float32)
        keys.append(k)
    for n in range(17, 64):
        base = u.sort_adversary(n)
        keys += [base.copy(), np.floor(base / 2).astype(np.float32)]
    for k in keys:
        perms.append(u.std_sort(k)[0])
    lens = np.array([len(k) for k in keys], np.int32)
    return dict(lens=lens, keys=np.concatenate(keys).astype(np.float32), perms=np.concatenate(perms).astype(np.int32))


def main():
    assert po.have_ref(), "oracle/_ref/libisx_ref.so missing: run `make -C oracle ref` where /root/reference exists"
    only = set(sys.argv[1:])                      # python make_golden.py [case ...]: regenerate just these fixtures
    if not only or "sort_vectors" in only:
        np.savez_compressed(os.path.join(HERE, "sort_vectors.npz"), **sort_vectors())
    for name, kw in CASES.items():
        if only and name not in only:
            continue
        out = run(po.RefEnv, kw)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "status histogram", np.bincount(out["status"].ravel().astype(np.int64), minlength=6).tolist())
    if only:
        return
    # geometry + route golden
    u = po.ref_unit()
    geo = {}
    for L in (2, 3):
        geo[f"road_{L}"] = np.packbits(u.road_map(L))
        geo[f"line_{L}"] = np.packbits(u.line_map(L))
        ids = [f"IN_{k}" for k in range(1, 4 * L + 1)] + [f"OUT_{k}" for k in range(1, 4 * L + 1)]
        paths, meta = [], []
        for a in ids:
            for b in ids:
                n, p, intent, sp = u.route(L, a, b)
                paths.append(p)
                meta.append([intent, *sp.view(np.uint32).tolist()])
        geo[f"paths_{L}"] = np.stack(paths).astype(np.float32)
        geo[f"meta_{L}"] = np.array(meta, np.int64)
    np.savez_compressed(os.path.join(HERE, "geometry_routes.npz"), **geo)
    # libm probe points: values of the libm entry points the reference binds, on this machine
    rng = np.random.default_rng(7)
    a = np.concatenate([rng.uniform(-7, 7, 20000), rng.uniform(-0.8, 0.8, 5000), rng.uniform(-900, 900, 5000)]).astype(np.float32)
    b = rng.uniform(-900, 900, a.size).astype(np.float32)
    s, c = u.sincosf(a)
    np.savez_compressed(os.path.join(HERE, "libm_points.npz"), a=a, b=b, sin=s, cos=c, tan=u.tanf(a), atan2=u.atan2f(a, b), hypot=u.hypotf(a, b))
    print("golden written to", HERE)


if __name__ == "__main__":
    main()
