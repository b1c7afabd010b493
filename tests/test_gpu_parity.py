"""GPU parity tests proper: the CUDA stepper (through the C ABI) against the CPU checker on identical seeds,
actions and routes.  Everything is compared BIT-EXACT: obs, rewards, flags, statuses, lidar hit indices, ego and
NPC state, NPC spawn/removal events (stricter than the 1e-5 relative tolerance BASELINE.json allows for floats)."""
import numpy as np
import pytest

import pyoracle as po
from parity_util import checker_class, free_run, make_pair, pursuit_policy

pytestmark = pytest.mark.gpu

R3 = po.ROUTES_3LANES
R2 = po.ROUTES_2LANES


def _benv():
    from marl_traffic_intersection_b200 import BatchedIntersectionEnv
    return BatchedIntersectionEnv


CONFIGS = {
    # BASELINE.json configs[0..4] shapes at parity-test sizes
    "C1_single": dict(num_envs=4, num_agents=1, num_lanes=3, ego_routes=[("IN_6", "OUT_2")]),
    "C2_team3": dict(num_envs=6, num_agents=3, num_lanes=3, ego_routes=[("IN_6", "OUT_2"), ("IN_4", "OUT_8"), ("IN_5", "OUT_7")],
                     use_team_reward=True),
    "C3_traffic": dict(num_envs=6, num_agents=1, num_lanes=3, ego_routes=[("IN_6", "OUT_2")], traffic_flow=True, traffic_density=0.5),
    "C4_eight": dict(num_envs=5, num_agents=8, num_lanes=3, ego_routes=R3[:8]),
    "C5_eight_traffic72": dict(num_envs=5, num_agents=8, num_lanes=3, ego_routes=R3[:8], traffic_flow=True, traffic_density=1.0,
                               lidar_rays=72),
}


@pytest.mark.parametrize("name", list(CONFIGS))
def test_free_running_random_actions(name):
    cfg = CONFIGS[name]
    b, refs = make_pair(_benv(), cfg, seed=0)
    counts = free_run(b, refs, steps=400, seed=0)
    assert counts.sum() == 400 * b.num_envs * b.num_agents
    b.close()


@pytest.mark.parametrize("seed", [1, 2])
def test_config2_2000_steps(seed):
    b, refs = make_pair(_benv(), CONFIGS["C2_team3"] | dict(num_envs=3), seed=seed)
    counts = free_run(b, refs, steps=2000, seed=seed)
    assert counts[po.STATUS_NAMES.index("CRASH_WALL")] + counts[po.STATUS_NAMES.index("CRASH_LINE")] > 0
    b.close()


def test_route_following_reaches_success_and_terminates():
    cfg = dict(num_envs=4, num_agents=2, num_lanes=3, ego_routes=[("IN_6", "OUT_2"), ("IN_1", "OUT_4")], max_steps=600)
    b, refs = make_pair(_benv(), cfg, seed=3)
    counts = free_run(b, refs, steps=900, seed=3, policy=pursuit_policy())
    assert counts[po.STATUS_NAMES.index("SUCCESS")] > 0
    b.close()


def test_dense_traffic_npc_collisions_and_yielding():
    cfg = dict(num_envs=6, num_agents=1, num_lanes=3, ego_routes=[("IN_6", "OUT_2")], traffic_flow=True, traffic_density=30.0,
               npc_capacity=32)
    b, refs = make_pair(_benv(), cfg, seed=5)
    free_run(b, refs, steps=700, seed=5, policy=pursuit_policy(throttle=0.2))
    st = b.stats()
    assert st["npc_spawned"] > 20 and st["npc_removed"] > 5 and st["npc_overflow"] == 0
    b.close()


def test_two_lanes_no_respawn_host_api():
    cfg = dict(num_envs=4, num_agents=4, num_lanes=2, ego_routes=R2[:4], traffic_flow=True, traffic_density=3.0,
               respawn_enabled=False, max_steps=300)
    b, refs = make_pair(_benv(), cfg, seed=7)
    free_run(b, refs, steps=500, seed=7, host_api=True)
    b.close()


def test_env_id_base_sharding_is_transparent():
    """Shard [4,8) of a larger job == envs 4..7 of the checker (RNG keyed by global env id)."""
    cfg = dict(num_envs=4, num_agents=1, num_lanes=3, ego_routes=[("IN_6", "OUT_2")], traffic_flow=True, traffic_density=2.0)
    b, refs = make_pair(_benv(), cfg, seed=11, env_id_base=4)
    free_run(b, refs, steps=300, seed=11, env_id_base=4)
    b.close()


def test_on_device_rollout_matches_checker_rollout():
    """isx_rollout (on-device Philox actions + auto-reset) against the checker's own rollout loop: status histogram
    identical, reward sum equal to double rounding of the same float32 addends."""
    cfg = dict(num_envs=8, num_agents=3, num_lanes=3, ego_routes=[("IN_6", "OUT_2"), ("IN_4", "OUT_8"), ("IN_5", "OUT_7")],
               use_team_reward=True, traffic_flow=True, traffic_density=1.0, max_steps=150, auto_reset=True)
    b, refs = make_pair(_benv(), cfg, seed=9)
    steps = 500
    b.rollout(steps)
    st = b.stats()
    hist = np.zeros(6, np.int64)
    rsum = 0.0
    for r in refs:
        n, h, rs = r.rollout(steps)
        hist += h
        rsum += rs
    got = np.array([st["status_hist"][k] for k in po.STATUS_NAMES])
    assert (got == hist).all(), (got, hist)
    assert st["agent_steps"] == steps * 8 * 3
    assert abs(st["reward_sum"] - rsum) <= 1e-9 * max(1.0, abs(rsum))
    b.close()


def test_state_injection_roundtrip_and_resync():
    """set_env_state / get_env_state (get_state / set_state of the reference) + one transition from an injected state."""
    cfg = dict(num_envs=2, num_agents=2, num_lanes=3, ego_routes=[("IN_6", "OUT_2"), ("IN_5", "OUT_7")], traffic_flow=True,
               traffic_density=2.0)
    b, refs = make_pair(_benv(), cfg, seed=13)
    free_run(b, refs, steps=120, seed=13)
    # copy env 0's state (checker) into env 1 of both sides, then step both and compare
    import ctypes as C
    from marl_traffic_intersection_b200 import _lib
    eg, npc = refs[0].egos(), refs[0].npcs()
    refs[1].set_egos(eg)
    refs[1].set_npcs(npc)
    refs[1].step_count = refs[0].step_count
    ce = (_lib.CarState * len(eg))()
    cn = (_lib.CarState * max(len(npc), 1))()
    for arr, src in ((ce, eg), (cn, npc)):
        for i, s in enumerate(src):
            for f in ("x", "y", "v", "heading", "acc", "steer", "prev_dist", "prev_a0", "prev_a1"):
                setattr(arr[i], f, float(s[f]))
            for f in ("path_index", "route", "alive", "uid", "intention"):
                setattr(arr[i], f, int(s[f]))
    b.set_env_state(1, ce, cn, len(npc), refs[0].step_count, refs[1].tick)
    g_eg, g_npc, n, sc, tk = b.get_env_state(1)
    assert n == len(npc) and sc == refs[0].step_count and tk == refs[1].tick
    assert all(np.float32(g_eg[i].x) == eg["x"][i] for i in range(len(eg)))
    import torch
    from parity_util import compare_step
    act = np.stack([po.philox_actions(13, e, refs[e].tick + 1, 2) for e in range(2)])
    b.step(torch.from_numpy(act).cuda())
    outs = [refs[e].step(act[e]) for e in range(2)]
    compare_step(b, refs, outs, "after injection")
    b.close()


def test_pipelined_host_step_equals_device_step():
    """isx_step_pinned cuts the env range into 4 shards and overlaps their device->host copies with the next shard's
    kernels; results must be identical to the single-launch device step (and therefore to the checker)."""
    import torch
    cfg = dict(num_envs=1024, num_agents=3, num_lanes=3, ego_routes=[("IN_6", "OUT_2"), ("IN_4", "OUT_8"), ("IN_5", "OUT_7")],
               use_team_reward=True, traffic_flow=True, traffic_density=2.0, seed=21)
    a_env, b_env = _benv()(cfg), _benv()(cfg)
    rng = np.random.default_rng(0)
    for t in range(60):
        act = rng.uniform(-1, 1, (1024, 3, 2)).astype(np.float32)
        obs_d, rew_d, term_d, trunc_d, info = a_env.step(torch.from_numpy(act).cuda())
        obs_h, rew_h, done_h, status_h, term_h, trunc_h = b_env.step_host(act)
        torch.cuda.synchronize()
        assert (obs_d.cpu().numpy().view(np.uint32) == obs_h.view(np.uint32)).all(), t
        assert (rew_d.cpu().numpy().view(np.uint32) == rew_h.view(np.uint32)).all(), t
        assert (info["status"].cpu().numpy() == status_h).all() and (info["done"].cpu().numpy() == done_h).all()
        assert (term_d.cpu().numpy() == term_h).all() and (trunc_d.cpu().numpy() == trunc_h).all()
    for k in ("npc_x", "npc_count", "ego_x", "lidar_hit", "events"):
        assert (a_env.buf[k].cpu().numpy() == b_env.buf[k].cpu().numpy()).all(), k
    a_env.close(); b_env.close()


def test_full_size_shard_invariance_and_spot_check():
    """BASELINE.json's full per-GPU size (8192 envs x 8 agents + traffic, 72 beams): (1) the job split into two
    4096-env shards with env_id_base gives the very same bits as the single 8192-env job (size-independent
    property: results do not depend on batching or sharding); (2) a few env ids spot-checked against the checker."""
    import torch
    base = dict(num_agents=8, num_lanes=3, ego_routes=R3[:8], traffic_flow=True, traffic_density=1.0, lidar_rays=72, seed=17,
                auto_reset=True, max_steps=40)
    full = _benv()(dict(base, num_envs=8192))
    lo = _benv()(dict(base, num_envs=4096, env_id_base=0))
    hi = _benv()(dict(base, num_envs=4096, env_id_base=4096))
    spots = [0, 1, 4095, 4096, 8191]
    cls = checker_class()
    refs = {e: cls(num_lanes=3, ego_routes=R3[:8], traffic=True, density=1.0, lidar_rays=72, seed=17, env_id=e, max_steps=40) for e in spots}
    for t in range(60):                       # crosses the max_steps=40 truncation -> auto-reset on every env
        full.rollout(1); lo.rollout(1); hi.rollout(1)
        torch.cuda.synchronize()
        for k in ("obs", "reward", "status", "terminated", "truncated", "lidar_hit", "npc_count", "npc_x", "ego_x", "ego_heading"):
            f = full.buf[k]
            assert torch.equal(f[:4096], lo.buf[k]) and torch.equal(f[4096:], hi.buf[k]), (t, k)
        obs = full.buf["obs"].cpu().numpy()
        st = full.buf["status"].cpu().numpy()
        for e, r in refs.items():
            a = po.philox_actions(17, e, r.tick + 1, 8)
            o = r.step(a)
            assert (obs[e].view(np.uint32) == o["obs"].view(np.uint32)).all(), (t, e)
            assert (st[e] == o["status"]).all(), (t, e)
            if o["terminated"] or o["truncated"]:
                r.reset()                     # the stepper auto-resets at the start of its next step
    s = full.stats()
    assert s["agent_steps"] == 8192 * 8 * 60 and s["env_resets"] == 8192 and s["npc_overflow"] == 0
    for x in (full, lo, hi):
        x.close()


def test_device_snapshot_restore_replays_bit_exactly():
    """isx_snapshot_save / _restore (the reference's get_state / set_state, for the whole batch, on the device): replaying
    the same actions from a restored snapshot reproduces every bit; a masked restore rolls back only the chosen envs."""
    import torch
    cfg = dict(num_envs=64, num_agents=3, num_lanes=3, ego_routes=[("IN_6", "OUT_2"), ("IN_4", "OUT_8"), ("IN_5", "OUT_7")],
               use_team_reward=True, traffic_flow=True, traffic_density=3.0, seed=31)
    b = _benv()(cfg)
    rng = np.random.default_rng(1)
    acts = [torch.from_numpy(rng.uniform(-1, 1, (64, 3, 2)).astype(np.float32)).cuda() for _ in range(80)]
    for t in range(40):
        b.step(acts[t])
    snap = b.snapshot()
    obs_at_save = b.buf["obs"].clone()
    first = []
    for t in range(40, 80):
        b.step(acts[t])
        first.append({k: b.buf[k].clone() for k in ("obs", "reward", "status", "npc_x", "npc_count", "lidar_hit", "ego_x")})
    b.restore(snap)
    assert torch.equal(b.buf["obs"], obs_at_save)
    for i, t in enumerate(range(40, 80)):
        b.step(acts[t])
        for k, v in first[i].items():
            assert torch.equal(b.buf[k], v), (t, k)
    # masked: roll back the even envs only, then one more common step; odd envs continue from step 80
    end_state = {k: b.buf[k].clone() for k in ("ego_x", "npc_count", "step")}
    mask = torch.zeros(64, dtype=torch.uint8, device="cuda"); mask[::2] = 1
    b.restore(snap, mask)
    assert torch.equal(b.buf["ego_x"][1::2], end_state["ego_x"][1::2]) and torch.equal(b.buf["step"][1::2], end_state["step"][1::2])
    assert (b.buf["step"][::2] == 40).all() and torch.equal(b.buf["obs"][::2], obs_at_save[::2])
    b.close()
