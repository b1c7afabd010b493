"""Round-2 additions to the C ABI, each against the CPU checker: the split statistics collective (int64 counters + float64
reward sum, 2-rank NCCL), next-step auto-reset (auto_reset = 2), run-time lidar beam count and live reconfiguration."""
import os
import sys

import numpy as np
import pytest

import pyoracle as po
from parity_util import bits, checker_class, compare_step, make_pair

pytestmark = pytest.mark.gpu
R3 = po.ROUTES_3LANES


def _benv():
    from marl_traffic_intersection_b200 import BatchedIntersectionEnv
    return BatchedIntersectionEnv


def test_stats_tensors_are_two_typed_views():
    import torch
    b = _benv()(dict(num_envs=300, num_agents=3, num_lanes=3, ego_routes=R3[:3], traffic_flow=True, traffic_density=2.0, auto_reset=True,
                     max_steps=50, seed=3))
    b.rollout(130)
    cnt, rs = b.stats_tensors()
    assert cnt.dtype == torch.int64 and cnt.shape == (15,) and rs.dtype == torch.float64 and rs.shape == (1,)
    st = b.stats()
    tot = b.reduce_stats()
    assert tot["agent_steps"] == st["agent_steps"] == 300 * 3 * 130 and tot["status_hist"] == st["status_hist"]
    assert tot["reward_sum"] == st["reward_sum"] and tot["reward_sum"] != 0.0 and tot["env_resets"] == st["env_resets"] == 600
    for k in ("npc_spawned", "npc_removed", "npc_collided", "npc_overflow", "neighbor_tie_sorts"):
        assert tot[k] == st[k]
    b.close()


def _nccl_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "oracle"), os.path.join(root, "tests")]
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from marl_traffic_intersection_b200 import BatchedIntersectionEnv
    per = 64
    b = BatchedIntersectionEnv(dict(num_envs=per, num_agents=3, num_lanes=3, ego_routes=R3[:3], traffic_flow=True, traffic_density=2.0,
                                    auto_reset=True, max_steps=60, seed=9, env_id_base=rank * per, device=f"cuda:{rank}"))
    b.rollout(150)
    local = b.stats()
    tot = b.reduce_stats()                       # the product's collective: NCCL all-reduce of int64[15] and of float64[1]
    gathered = [None] * world
    dist.all_gather_object(gathered, (local["reward_sum"], local["agent_steps"], local["status_hist"]))
    if rank == 0:
        q.put((tot, gathered))
    dist.barrier()
    b.close()
    dist.destroy_process_group()


def test_two_rank_nccl_reduction_equals_unsharded_job():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1500)
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    tot, gathered = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert tot["reward_sum"] == gathered[0][0] + gathered[1][0]              # sum of the per-rank DOUBLES
    assert tot["agent_steps"] == gathered[0][1] + gathered[1][1] == 2 * 64 * 3 * 150
    whole = _benv()(dict(num_envs=128, num_agents=3, num_lanes=3, ego_routes=R3[:3], traffic_flow=True, traffic_density=2.0,
                         auto_reset=True, max_steps=60, seed=9))
    whole.rollout(150)
    w = whole.stats()
    whole.close()
    for k in ("agent_steps", "status_hist", "npc_spawned", "npc_removed", "npc_collided", "env_resets"):
        assert tot[k] == w[k], k                                               # counters == the unsharded job
    assert abs(tot["reward_sum"] - w["reward_sum"]) <= 1e-9 * max(1.0, abs(w["reward_sum"]))


def test_next_step_auto_reset_returns_the_reset_observation():
    """auto_reset = 2: the call after a terminated|truncated step only resets — action ignored, reward 0, flags 0, step 0,
    RNG tick unchanged — and returns what reset() returns; afterwards the env keeps matching a checker that was reset."""
    import torch
    cfg = dict(num_envs=6, num_agents=3, num_lanes=3, ego_routes=R3[:3], use_team_reward=True, traffic_flow=True, traffic_density=3.0,
               max_steps=35, auto_reset=2, respawn_enabled=False)
    b, refs = make_pair(_benv(), cfg, seed=23)
    b.reset()
    for r in refs:
        r.reset()
    pending = np.zeros(6, bool)
    resets = 0
    for t in range(260):
        act = np.stack([po.philox_actions(23, e, refs[e].tick + 1, 3) for e in range(6)])
        b.step(torch.from_numpy(act).cuda())
        torch.cuda.synchronize()
        buf = {k: v.cpu().numpy() for k, v in b.buf.items()}
        nxt = np.zeros(6, bool)
        for e in range(6):
            if pending[e]:
                refs[e].reset()                                    # what the library did instead of stepping
                resets += 1
                assert (bits(buf["obs"][e]) == bits(refs[e].obs())).all(), (t, e)
                assert (buf["reward"][e] == 0).all() and (buf["done"][e] == 0).all() and (buf["status"][e] == 0).all()
                assert not buf["terminated"][e] and not buf["truncated"][e] and buf["step"][e] == 0 and buf["npc_count"][e] == 0
                assert int(buf["tick"][e]) == refs[e].tick
            else:
                o = refs[e].step(act[e])
                compare_step(_One(b, e), [refs[e]], [o], f"step {t} env {e}")
                nxt[e] = o["terminated"] or o["truncated"]
        pending = nxt
    assert resets >= 12
    st = b.stats()
    assert st["env_resets"] == resets and st["agent_steps"] == (260 * 6 - resets) * 3
    b.close()


class _One:
    """View of one env of a batch with the attributes compare_step reads."""

    def __init__(self, b, e):
        self.num_envs, self.num_agents, self.traffic_flow = 1, b.num_agents, b.traffic_flow
        self.buf = {k: v[e:e + 1] for k, v in b.buf.items()}


def test_runtime_lidar_rays_and_live_settings():
    """isx_set_lidar_rays / isx_set_reward / isx_configure_episode / isx_set_traffic_density on a live handle == a checker
    reconfigured the same way (Lidar objects swapped, reward_config fields written, configure*/ called again)."""
    import torch
    cfg = dict(num_envs=4, num_agents=3, num_lanes=3, ego_routes=R3[:3], traffic_flow=True, traffic_density=1.0, lidar_rays=96)
    b, refs = make_pair(_benv(), cfg, seed=29)
    b.reset()
    for r in refs:
        r.reset()

    def run(n, tag):
        for t in range(n):
            act = np.stack([po.philox_actions(29, e, refs[e].tick + 1, 3) for e in range(4)])
            b.step(torch.from_numpy(act).cuda())
            outs = [refs[e].step(act[e]) for e in range(4)]
            compare_step(b, refs, outs, f"{tag} {t}")
    run(60, "96 beams")
    b.set_lidar_rays(72)
    for r in refs:
        r._f("set_lidar_rays")(r._h, 72)
        r.lidar_rays = 72
        if isinstance(r, po.RefEnv):
            r._lib.isxref_swap_lidars.argtypes = [__import__("ctypes").c_void_p]
            r._lib.isxref_swap_lidars(r._h)                         # default-constructed 72-beam Lidar objects, as set_state leaves them
    torch.cuda.synchronize()
    o = b.buf["obs"].cpu().numpy()
    assert (o[:, :, 31:103] == 1.0).all() and (o[:, :, 103:] == 0.0).all()
    run(60, "72 beams")
    new_reward = (4.0, 2.0, -0.3, -6.0, -2.0, 15.0, -0.1, 0.4)
    b.set_reward_config(new_reward)
    b.configure(True, False, 500)
    b.set_traffic_density(6.0)
    for r in refs:
        k = np.array(new_reward, np.float32)
        r._f("set_reward")(r._h, k.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_float)))
        r._f("configure")(r._h, 1, 0, 500)
        r._f("configure_traffic")(r._h, 1, 6.0)
    run(80, "reconfigured")
    b.close()


@pytest.mark.parametrize("shape", ["odd_sizes_traffic", "max_agents_generic_rays", "two_lanes_96"])
def test_no_kernel_stores_outside_its_buffers(shape, monkeypatch):
    """compute-sanitizer is not available on this GPU pool, so the library brings its own check: with ISX_GUARD=1 every
    device buffer sits between red zones; after exercising every kernel and every mode on awkward sizes (env counts that
    fill no warp, the maximum ego and NPC counts, beam counts that fill no piece, masked reset, snapshot / restore,
    observe, both transports of the host step) no red zone may have been written to."""
    import torch
    monkeypatch.setenv("ISX_GUARD", "1")
    R3 = po.ROUTES_3LANES
    if shape == "odd_sizes_traffic":
        cfgs = [dict(num_envs=E, num_agents=3, ego_routes=R3[:3], use_team_reward=True, traffic_flow=True, traffic_density=20.0,
                     lidar_rays=rays, seed=1, auto_reset=ar, max_steps=25, npc_capacity=cap)
                for E, rays, ar, cap in ((5, 72, 1, 16), (1030, 96, 2, 16), (6151, 72, 1, 8), (3, 7, 1, 32))]
    elif shape == "max_agents_generic_rays":
        cfgs = [dict(num_envs=7, num_agents=32, ego_routes=[R3[i % 12] for i in range(32)], traffic_flow=True, traffic_density=30.0,
                     lidar_rays=rays, seed=2, auto_reset=1, max_steps=40, npc_capacity=32) for rays in (33, 95, 1)]
    else:
        cfgs = [dict(num_envs=129, num_agents=5, num_lanes=2, ego_routes=po.default_routes(2)[:5], traffic_flow=True, traffic_density=5.0,
                     lidar_rays=96, seed=3, auto_reset=1, max_steps=30)]
    for cfg in cfgs:
        for lanes in ("8", "16", "32"):
            monkeypatch.setenv("ISX_TRAFFIC_LANES", lanes)
            b = _benv()(dict(cfg))
            E, N = b.num_envs, b.num_agents
            b.reset()
            b.rollout(35)
            snap = b.snapshot()
            a = torch.rand(E, N, 2, device="cuda") * 2 - 1
            for _ in range(4):
                b.step(a)
            m = torch.zeros(E, dtype=torch.uint8, device="cuda")
            m[::2] = 1
            b.reset(m); b.restore(snap, m); b.restore(snap); b.observe()
            b.step_host(np.zeros((E, N, 2), np.float32))
            b.set_lidar_rays(max(1, b.lidar_rays - 1))
            b.rollout(3)
            b.render(0)
            b.stats()
            assert b.check_guards() == 0, (cfg, lanes)
            b.close()


def test_guard_bands_do_detect_a_stray_store(monkeypatch):
    """The detector itself: a 4-byte store just past the end of one buffer (issued from here with cudaMemset) is reported."""
    from cuda import cudart
    monkeypatch.setenv("ISX_GUARD", "1")
    b = _benv()(dict(num_envs=3, num_agents=2, ego_routes=po.ROUTES_3LANES[:2]))
    b.reset()
    b.rollout(2)
    assert b.check_guards() == 0
    r = b.buf["obs"]                          # its own allocation (reward / done / status share one block)
    (err,) = cudart.cudaMemset(r.data_ptr() + r.numel() * r.element_size(), 0, 4)
    assert int(err) == 0
    assert b.check_guards() == 1
    (err,) = cudart.cudaMemset(b.buf["ego_x"].data_ptr() - 8, 0, 1)
    assert int(err) == 0 and b.check_guards() == 2
    b.close()
    monkeypatch.delenv("ISX_GUARD")
    b = _benv()(dict(num_envs=3, num_agents=2, ego_routes=po.ROUTES_3LANES[:2]))
    with pytest.raises(Exception):
        b.check_guards()                       # guards are off: ISX_E_STATE
    b.close()


@pytest.mark.parametrize("E", [640, 6144])
def test_step_host_into_caller_buffers_equals_pinned_views(E):
    """isx_step_host with caller-owned (pageable) output buffers — what a ctypes / cgo binding passes — against the same step
    read from the library's pinned views, for both transports (rows copied by the copy engine below 16,384 agents, compact
    records expanded by host threads above)."""
    import ctypes as C
    from marl_traffic_intersection_b200 import _lib
    cfg = dict(num_envs=E, num_agents=3, ego_routes=po.ROUTES_3LANES[:3], traffic_flow=True, traffic_density=2.0, seed=9)
    a_env, b_env = _benv()(cfg), _benv()(cfg)
    rng = np.random.default_rng(3)
    obs = np.full((E, 3, 127), np.nan, np.float32)
    rew = np.zeros((E, 3), np.float32)
    done, status = np.zeros((E, 3), np.uint8), np.zeros((E, 3), np.uint8)
    term, trunc = np.zeros(E, np.uint8), np.zeros(E, np.uint8)
    vp = lambda x: C.c_void_p(x.ctypes.data)
    for t in range(12):
        act = rng.uniform(-1, 1, (E, 3, 2)).astype(np.float32)
        o1, r1, d1, s1, t1, u1 = a_env.step_host(act)
        _lib.check(b_env._lib, b_env._lib.isx_step_host(b_env._h, vp(act), C.c_float(1.0 / 60.0), vp(obs), vp(rew), vp(done), vp(status),
                                                      vp(term), vp(trunc), b_env._stream()))
        assert (bits(o1) == bits(obs)).all() and (bits(r1) == bits(rew)).all(), t
        assert (d1 == done).all() and (s1 == status).all() and (t1 == term.astype(bool)).all() and (u1 == trunc.astype(bool)).all(), t
    a_env.close(); b_env.close()


@pytest.mark.parametrize("lanes", ["8", "16"])
def test_traffic_env_order_changes_no_result(lanes, monkeypatch):
    """k_traffic<8 / 16> takes its envs from lists filed by NPC count (k_traffic_order) instead of in index order.  The
    same batches — two groups with different settings, the packed instance forced (the library picks it above 12,288 envs),
    next-step auto-reset, a masked reset, a host-buffer step cut into pipeline ranges — stepped with the lists
    (default) and without (ISX_NO_ORDER=1) must end bit-identical in every buffer and counter."""
    import torch
    cfgs = [dict(num_envs=5003, num_agents=2, num_lanes=3, ego_routes=R3[:2], traffic_flow=True, traffic_density=6.0, seed=4,
                 auto_reset=2, max_steps=90, npc_capacity=16),
            dict(num_envs=9001, num_agents=2, num_lanes=3, ego_routes=R3[2:4], traffic_flow=True, traffic_density=1.0, seed=5,
                 auto_reset=1, max_steps=70, npc_capacity=16)]
    ends = []
    monkeypatch.setenv("ISX_TRAFFIC_LANES", lanes)
    for no_order in (False, True):
        if no_order:
            monkeypatch.setenv("ISX_NO_ORDER", "1")
        else:
            monkeypatch.delenv("ISX_NO_ORDER", raising=False)
        b = _benv()([dict(c) for c in cfgs])
        E, N = b.num_envs, b.num_agents
        b.reset()
        b.rollout(120)
        m = torch.zeros(E, dtype=torch.uint8, device="cuda")
        m[::3] = 1
        b.reset(m)
        g = torch.Generator(device="cuda").manual_seed(1)
        for _ in range(5):
            b.step(torch.rand(E, N, 2, device="cuda", generator=g) * 2 - 1)
        b.step_host(np.full((E, N, 2), 0.25, np.float32))
        b.rollout(60)
        torch.cuda.synchronize()
        ends.append(({k: b.buf[k].cpu().numpy().copy() for k in ("obs", "reward", "status", "npc_count", "events", "lidar_hit", "terminated", "truncated")},
                     b.stats()))
        b.close()
    (a, sa), (c, sc) = ends
    assert sa == sc and sa["npc_spawned"] > 1000 and sa["env_resets"] > 1000
    for k in a:
        assert np.array_equal(a[k].view(np.uint8), c[k].view(np.uint8)), k
