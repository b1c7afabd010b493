"""GPU parity tests proper: the CUDA stepper (through the C ABI) against the CPU checker on identical seeds,
actions and routes.  Everything is compared BIT-EXACT: obs, rewards, flags, statuses, lidar hit indices, ego and
NPC state, NPC spawn/removal events (stricter than the 1e-5 relative tolerance BASELINE.json allows for floats)."""
import numpy as np
import pytest

import pyoracle as po
from parity_util import checker_class, compare_step, free_run, make_pair, pursuit_policy

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


@pytest.mark.parametrize("lanes", [8, 16, 32])
def test_dense_traffic_npc_collisions_and_yielding(lanes, monkeypatch):
    """Every k_traffic instance (8 / 16 / 32 lanes per env; the library picks by batch size, ISX_TRAFFIC_LANES forces one) on
    the same dense traffic: up to 13 NPCs per env, so the 8- and 16-lane instances also run their wide in-kernel fallback."""
    monkeypatch.setenv("ISX_TRAFFIC_LANES", str(lanes))
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


@pytest.mark.parametrize("E", [1024, 6144])
def test_pipelined_host_step_equals_device_step(E):
    """Both transports of the host-buffer step: 1024 envs x 3 egos ship the obs rows themselves (small batch, no host
    threads), 6144 x 3 ship compact records that the library's host threads expand.  isx_step_pinned cuts the env range into 4 shards and overlaps their device->host copies with the next shard's
    kernels, replaying the whole step from a CUDA graph captured per dt; results must be identical to the single-launch
    device step (and therefore to the checker), also when dt changes between calls (graph re-captured) and when device
    steps are interleaved on the caller's stream."""
    import torch
    cfg = dict(num_envs=E, num_agents=3, num_lanes=3, ego_routes=[("IN_6", "OUT_2"), ("IN_4", "OUT_8"), ("IN_5", "OUT_7")],
               use_team_reward=True, traffic_flow=True, traffic_density=2.0, seed=21)
    a_env, b_env = _benv()(cfg), _benv()(cfg)
    assert (b_env.host_step_bytes()["host_expand_threads"] > 0) == (E * 3 >= 16384)
    rng = np.random.default_rng(0)
    for t in range(60):
        act = rng.uniform(-1, 1, (E, 3, 2)).astype(np.float32)
        dt = (1.0 / 60.0, 1.0 / 60.0, 1.0 / 30.0, 0.02)[(t // 5) % 4]
        obs_d, rew_d, term_d, trunc_d, info = a_env.step(torch.from_numpy(act).cuda(), dt)
        if t % 7 == 3:                     # a device-buffer step in between: the replay must stay ordered behind it
            b_env.step(torch.from_numpy(act).cuda(), dt)
            act = -act
            obs_d, rew_d, term_d, trunc_d, info = a_env.step(torch.from_numpy(act).cuda(), dt)
        if t % 2:
            obs_h, rew_h, done_h, status_h, term_h, trunc_h = b_env.step_host(act, dt)
        else:                                 # in-place variant: actions written straight into the pinned staging buffer
            b_env.host_actions[...] = act
            obs_h, rew_h, done_h, status_h, term_h, trunc_h = b_env.step_host(None, dt)
        torch.cuda.synchronize()
        assert (obs_d.cpu().numpy().view(np.uint32) == obs_h.view(np.uint32)).all(), t
        assert (rew_d.cpu().numpy().view(np.uint32) == rew_h.view(np.uint32)).all(), t
        assert (info["status"].cpu().numpy() == status_h).all() and (info["done"].cpu().numpy() == done_h).all()
        assert (term_d.cpu().numpy() == term_h).all() and (trunc_d.cpu().numpy() == trunc_h).all()
    for k in ("npc_x", "npc_count", "ego_x", "lidar_hit", "events"):
        assert (a_env.buf[k].cpu().numpy() == b_env.buf[k].cpu().numpy()).all(), k
    a_env.close(); b_env.close()


@pytest.mark.parametrize("E", [8192, 65536])
def test_full_size_shard_invariance_and_spot_check(E):
    """BASELINE.json's full size — all 65,536 envs x 8 agents + traffic, 72 beams on one GPU, and the 8192-env share of the
    8-GPU split: (1) the job cut into two half-size shards with env_id_base gives the very same bits as the single job
    (size-independent property: results do not depend on batching or sharding); (2) a few env ids spot-checked against
    the checker."""
    import torch
    H = E // 2
    base = dict(num_agents=8, num_lanes=3, ego_routes=R3[:8], traffic_flow=True, traffic_density=1.0, lidar_rays=72, seed=17,
                auto_reset=True, max_steps=40)
    full = _benv()(dict(base, num_envs=E))
    lo = _benv()(dict(base, num_envs=H, env_id_base=0))
    hi = _benv()(dict(base, num_envs=H, env_id_base=H))
    spots = [0, 1, H - 1, H, E - 1]
    cls = checker_class()
    refs = {e: cls(num_lanes=3, ego_routes=R3[:8], traffic=True, density=1.0, lidar_rays=72, seed=17, env_id=e, max_steps=40) for e in spots}
    spot_idx = torch.tensor(spots, device="cuda")
    for t in range(60):                       # crosses the max_steps=40 truncation -> auto-reset on every env
        full.rollout(1); lo.rollout(1); hi.rollout(1)
        torch.cuda.synchronize()
        for k in ("obs", "reward", "status", "terminated", "truncated", "lidar_hit", "npc_count", "npc_x", "ego_x", "ego_heading"):
            f = full.buf[k]
            assert torch.equal(f[:H], lo.buf[k]) and torch.equal(f[H:], hi.buf[k]), (t, k)
        obs = full.buf["obs"][spot_idx].cpu().numpy()
        st = full.buf["status"][spot_idx].cpu().numpy()
        for i, (e, r) in enumerate(refs.items()):
            a = po.philox_actions(17, e, r.tick + 1, 8)
            o = r.step(a)
            assert (obs[i].view(np.uint32) == o["obs"].view(np.uint32)).all(), (t, e)
            assert (st[i] == o["status"]).all(), (t, e)
            if o["terminated"] or o["truncated"]:
                r.reset()                     # the stepper auto-resets at the start of its next step
    s = full.stats()
    assert s["agent_steps"] == E * 8 * 60 and s["env_resets"] == E and s["npc_overflow"] == 0
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


EDGE = {
    # sub-warp layouts of k_ego (NP = 1, 2, 4, 8 with padding, 16, 32) and generic beam counts
    "n2": dict(num_envs=5, num_agents=2, num_lanes=3, ego_routes=R3[:2]),
    "n5_pad8": dict(num_envs=5, num_agents=5, num_lanes=3, ego_routes=R3[:5], use_team_reward=True, traffic_flow=True, traffic_density=2.0),
    "n12_np16": dict(num_envs=3, num_agents=12, num_lanes=3, ego_routes=R3, traffic_flow=True, traffic_density=1.0, lidar_rays=72),
    "rays_33_generic": dict(num_envs=4, num_agents=3, num_lanes=3, ego_routes=R3[3:6], lidar_rays=72, traffic_flow=True, traffic_density=4.0, max_steps=0),
    "lanes4": dict(num_envs=4, num_agents=4, num_lanes=4, ego_routes=[("IN_1", "OUT_9"), ("IN_6", "OUT_15"), ("IN_11", "OUT_2"), ("IN_16", "OUT_4")],
                   traffic_flow=True, traffic_density=3.0, traffic_routes=[("IN_2", "OUT_10"), ("IN_7", "OUT_13"), ("IN_12", "OUT_3"), ("IN_13", "OUT_7")]),
    "lanes1": dict(num_envs=4, num_agents=2, num_lanes=1, ego_routes=[("IN_1", "OUT_3"), ("IN_2", "OUT_1")], traffic_flow=True, traffic_density=3.0,
                   traffic_routes=[("IN_3", "OUT_1"), ("IN_4", "OUT_2")]),
    "norespawn_custom_reward": dict(num_envs=6, num_agents=3, num_lanes=3, ego_routes=R3[:3], respawn_enabled=False, max_steps=90, use_team_reward=True,
                                    reward_config=dict(progress_scale=3.0, stuck_speed_threshold=2.5, stuck_penalty=-0.5, crash_vehicle_penalty=-7.0,
                                                       crash_object_penalty=-3.0, success_reward=20.0, action_smoothness_scale=-0.3, team_alpha=0.7)),
}


@pytest.mark.parametrize("name", list(EDGE))
def test_edge_configurations(name):
    b, refs = make_pair(_benv(), EDGE[name], seed=41)
    free_run(b, refs, steps=260, seed=41)
    b.close()


@pytest.mark.parametrize("dt", [1.0 / 30.0, 0.0, 0.05])
def test_per_call_dt(dt):
    cfg = dict(num_envs=4, num_agents=2, num_lanes=3, ego_routes=R3[4:6], traffic_flow=True, traffic_density=5.0)
    b, refs = make_pair(_benv(), cfg, seed=43)
    free_run(b, refs, steps=200, seed=43, dt=dt)
    b.close()


def test_out_of_range_actions_and_dead_egos():
    """|steer| > 1 drives tanf through its range-reduction branch (Car.cpp:14,28); an ego whose `alive` flag was cleared
    by the caller (set_state) reports DEAD, done=1, reward 0 and an all-zero obs row (IntersectionEnv.cpp:167-171,426-429)."""
    import torch
    from marl_traffic_intersection_b200 import _lib
    cfg = dict(num_envs=3, num_agents=3, num_lanes=3, ego_routes=R3[:3], use_team_reward=True)
    b, refs = make_pair(_benv(), cfg, seed=47)

    def wild(obs, t):
        rng = np.random.default_rng(t)
        return rng.uniform(-6, 6, (3, 3, 2)).astype(np.float32)
    free_run(b, refs, steps=120, seed=47, policy=wild)
    # kill ego 1 of env 2 on both sides
    eg = refs[2].egos()
    eg["alive"][1] = 0
    refs[2].set_egos(eg)
    ce = (_lib.CarState * 3)()
    for i, s in enumerate(eg):
        for f in ("x", "y", "v", "heading", "acc", "steer", "prev_dist", "prev_a0", "prev_a1"):
            setattr(ce[i], f, float(s[f]))
        for f in ("path_index", "route", "alive", "uid", "intention"):
            setattr(ce[i], f, int(s[f]))
    b.set_env_state(2, ce, None, 0, refs[2].step_count, refs[2].tick)
    from parity_util import compare_step
    for t in range(40):
        act = np.stack([po.philox_actions(47, e, refs[e].tick + 1, 3) for e in range(3)])
        b.step(torch.from_numpy(act).cuda())
        outs = [refs[e].step(act[e]) for e in range(3)]
        compare_step(b, refs, outs, f"dead-ego step {t}")
        assert outs[2]["status"][1] == po.STATUS_NAMES.index("DEAD") and (outs[2]["obs"][1] == 0).all()
    b.close()


def test_npc_capacity_overflow_is_counted_not_fatal():
    b = _benv()(dict(num_envs=64, num_agents=1, num_lanes=3, ego_routes=[("IN_6", "OUT_2")], traffic_flow=True, traffic_density=200.0,
                     npc_capacity=3, seed=5, auto_reset=True))
    b.rollout(400)
    st = b.stats()
    assert st["npc_overflow"] > 0 and int(b.buf["npc_count"].max()) <= 3 and st["agent_steps"] == 64 * 400
    b.close()


def test_full_eight_lane_group_matches_until_the_first_dropped_spawn(monkeypatch):
    """k_traffic<8> gives an env 8 lanes (four envs per warp; forced here, small batches would get a warp per env).  With npc_capacity = 8 there is no wide fallback: a group whose
    8 lanes all hold an NPC must still match the reference bit for bit, up to the step at which the bounded list drops a
    spawn the reference's unbounded list would take (counted in npc_overflow; from there the two legitimately differ)."""
    import torch
    monkeypatch.setenv("ISX_TRAFFIC_LANES", "8")
    cfg = dict(num_envs=8, num_agents=1, num_lanes=3, ego_routes=[("IN_6", "OUT_2")], traffic_flow=True, traffic_density=30.0,
               npc_capacity=8)
    b, refs = make_pair(_benv(), cfg, seed=11)
    b.reset()
    for r in refs:
        r.reset()
    full, t = False, 0
    for t in range(1, 1500):
        act = np.stack([po.philox_actions(11, e, refs[e].tick + 1, 1) for e in range(8)])
        b.step(torch.from_numpy(act).cuda())
        if b.stats()["npc_overflow"] > 0:
            break
        outs = [refs[e].step(act[e]) for e in range(8)]
        buf = compare_step(b, refs, outs, f"step {t}")
        full = full or bool((buf["npc_count"] == 8).any())
        need = np.array([o["terminated"] or o["truncated"] for o in outs])
        if need.any():
            b.reset(torch.from_numpy(need.astype(np.uint8)).cuda())
            for e in np.nonzero(need)[0]:
                refs[e].reset()
    assert full and t > 30, (full, t)
    b.close()


@pytest.mark.parametrize("n,cap,density", [(20, 0, 0.0), (20, 16, 3.0), (32, 32, 6.0)])
def test_equal_distance_neighbours_follow_std_sort(n, cap, density):
    """More egos than spawn lanes: several egos sit on the same spawn point, so neighbour distances tie exactly while the
    list is longer than 16 — the regime where IntersectionEnv.cpp:490's std::sort is not stable and the outcome is
    libstdc++'s introsort.  The device replays that order (k_features -> exact_neighbor_top5) and counts how often."""
    cfg = dict(num_envs=3, num_agents=n, num_lanes=3, ego_routes=[R3[i % 12] for i in range(n)], traffic_flow=density > 0,
               traffic_density=density, npc_capacity=cap, max_steps=120)
    b, refs = make_pair(_benv(), cfg, seed=53)
    free_run(b, refs, steps=200, seed=53)
    assert b.stats()["neighbor_tie_sorts"] > 0
    b.close()


HETERO = [
    dict(num_envs=3, num_agents=4, num_lanes=2, ego_routes=R2[:4], traffic_flow=True, traffic_density=3.0, seed=61, env_id_base=100, max_steps=150),
    dict(num_envs=2, num_agents=4, num_lanes=3, ego_routes=R3[:4], use_team_reward=True, seed=62, env_id_base=7, respawn_enabled=False, max_steps=90),
    dict(num_envs=4, num_agents=4, num_lanes=3, ego_routes=R3[4:8], traffic_flow=True, traffic_density=6.0, seed=63, env_id_base=0,
         reward_config=dict(progress_scale=2.0, crash_vehicle_penalty=-4.0, success_reward=30.0, team_alpha=0.5), use_team_reward=True),
    dict(num_envs=1, num_agents=4, num_lanes=4, ego_routes=[("IN_1", "OUT_9"), ("IN_6", "OUT_15"), ("IN_11", "OUT_2"), ("IN_16", "OUT_4")],
         traffic_flow=True, traffic_density=0.5, traffic_routes=[("IN_2", "OUT_10"), ("IN_7", "OUT_13")], seed=64, env_id_base=5),
]


def test_heterogeneous_batch_matches_per_env_checkers():
    """lanes / routes / traffic density / reward weights / episode settings differ per group inside ONE
    batch; every env must still match a checker env configured like its group — device step and host-buffer step."""
    from parity_util import make_group_pair
    for host_api in (False, True):
        b, refs = make_group_pair(_benv(), HETERO)
        assert b.group_ranges == [(0, 3), (3, 2), (5, 4), (9, 1)] and b.num_envs == 10

        def pol(obs, t):
            rng = np.random.default_rng(1000 + t)
            return rng.uniform(-1, 1, (10, 4, 2)).astype(np.float32)
        free_run(b, refs, steps=220, seed=0, policy=pol, host_api=host_api)
        b.close()


def test_heterogeneous_group_equals_standalone_batch():
    """Each group keeps its own seed / env_id_base, so after an on-device rollout (Philox actions, auto-reset) its slice of
    the heterogeneous batch is bit-identical to a stand-alone batch created from the same config; snapshots and
    per-env state access address envs by their position in the whole batch."""
    import torch
    cfgs = [dict(c, auto_reset=True, max_steps=60) for c in HETERO]
    het = _benv()(cfgs)
    solo = [_benv()(c) for c in cfgs]
    snap = het.snapshot()
    het.rollout(150)
    for s in solo:
        s.rollout(150)
    torch.cuda.synchronize()
    for (first, cnt), s in zip(het.group_ranges, solo):
        for k in ("obs", "reward", "status", "ego_x", "ego_heading", "npc_count", "tick", "step", "lidar_hit", "events"):
            a, c = het.buf[k][first:first + cnt].cpu().numpy(), s.buf[k].cpu().numpy()
            assert (a.view(np.uint8) == c.view(np.uint8)).all(), k
        n = int(s.buf["npc_count"].max())
        if n:
            assert (het.buf["npc_x"][first:first + cnt, :n].cpu().numpy().view(np.uint32) == s.buf["npc_x"][:, :n].cpu().numpy().view(np.uint32)).all()
    # per-env state access resolves the group (env 9 = the 4-lane group: intents come from ITS routes)
    egos, npcs, n_npc, sc, tk = het.get_env_state(9)
    e2, _, _, sc2, tk2 = solo[3].get_env_state(0)
    assert [c.intention for c in egos] == [c.intention for c in e2] and sc == sc2 and tk == tk2
    # masked restore across group boundaries: envs 2..6 go back to the reset state, the others keep going
    mask = torch.zeros(10, dtype=torch.uint8, device="cuda"); mask[2:7] = 1
    keep = het.buf["ego_x"].clone()
    het.restore(snap, mask)
    torch.cuda.synchronize()
    assert (het.buf["tick"][2:7] == 0).all() and (het.buf["ego_x"][:2] == keep[:2]).all() and (het.buf["ego_x"][7:] == keep[7:]).all()
    het.rollout(50)
    st = het.stats()
    assert st["agent_steps"] == 10 * 4 * 200
    het.close()
    for s in solo:
        s.close()


def test_heterogeneous_batch_rejects_shape_mismatch():
    with pytest.raises(Exception):
        _benv()([dict(num_envs=2, num_agents=2, ego_routes=R3[:2]), dict(num_envs=2, num_agents=3, ego_routes=R3[:3])])
    with pytest.raises(Exception):
        _benv()([dict(num_envs=2, num_agents=2, ego_routes=R3[:2], lidar_rays=72), dict(num_envs=2, num_agents=2, ego_routes=R3[:2], lidar_rays=96)])


def test_run_to_run_determinism_at_full_size():
    """Two batches created from the same config end a 300-step on-device rollout with identical bits everywhere (work is
    claimed dynamically from atomic counters in k_lidar_obs and kernels overlap through programmatic dependent launch:
    none of that may leak into results), and the host-buffer path agrees with them."""
    import torch
    cfg = dict(num_envs=4096, num_agents=8, num_lanes=3, ego_routes=R3[:8], traffic_flow=True, traffic_density=2.0, lidar_rays=72,
               seed=77, auto_reset=True, max_steps=120)
    a, b = _benv()(cfg), _benv()(cfg)
    a.rollout(300)
    for _ in range(3):
        b.rollout(100)
    torch.cuda.synchronize()
    for k in a.buf:
        x, y = a.buf[k].cpu().numpy(), b.buf[k].cpu().numpy()
        if k.startswith("npc_") and k != "npc_count":
            n = a.buf["npc_count"].cpu().numpy()
            m = np.arange(x.shape[1])[None, :] < n[:, None]
            x, y = np.where(m, x, 0), np.where(m, y, 0)
        assert (x.view(np.uint8) == y.view(np.uint8)).all(), k
    sa, sb = a.stats(), b.stats()
    assert sa == sb and sa["agent_steps"] == 4096 * 8 * 300
    a.close(); b.close()
