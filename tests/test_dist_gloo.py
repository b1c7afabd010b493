"""N>1 host logic on CPU (gloo, world size 2): env-id sharding + the one collective (sum of the episode counters).
The per-rank "device" is stood in for by the C oracle; what is under test is that shard [r*E/G, (r+1)*E/G) of a job,
reduced over ranks, equals the unsharded job — the property that lets bench.py scale with no per-step collective."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pyoracle as po

pytestmark = pytest.mark.skipif(not po.have_oracle(), reason="C oracle not built")
ROUTES = [("IN_6", "OUT_2"), ("IN_4", "OUT_8")]
STEPS, GLOBAL_ENVS, SEED = 300, 6, 5


def shard(rank, world, total):
    per = total // world
    return range(rank * per, (rank + 1) * per)


def counters_for(env_ids):
    v = np.zeros(15, np.int64)                      # layout of isx_stats_device_ptrs: int64[15], hist[0:6] ... agent_steps[11]
    rsum = 0.0
    for g in env_ids:
        e = po.OracleEnv(3, ROUTES, traffic=True, density=2.0, max_steps=120, seed=SEED, env_id=g)
        n, hist, rs = e.rollout(STEPS)
        v[0:6] += hist
        v[11] += n
        rsum += rs
    return v, rsum


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    v, rsum = counters_for(shard(rank, world, GLOBAL_ENVS))
    # the product's own reduction (BatchedIntersectionEnv.reduce_stats -> reduce_stat_tensors): int64 counters and the
    # float64 reward sum are reduced as two tensors — summing a double's bit pattern as an integer would be garbage
    from marl_traffic_intersection_b200.batched import reduce_stat_tensors
    t = torch.from_numpy(v)
    r = torch.tensor([rsum], dtype=torch.float64)
    tot = reduce_stat_tensors(t, r)
    assert tot["agent_steps"] == int(t[11]) and tot["reward_sum"] == float(r[0])
    try:
        reduce_stat_tensors(torch.zeros(16, dtype=torch.int64), torch.zeros(1, dtype=torch.int64))
        raise SystemExit("a reward sum carried as int64 must be refused")
    except TypeError:
        pass
    tmax = torch.tensor([float(rank + 1)], dtype=torch.float64)   # bench.py takes the MAX time over ranks
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    if rank == 0:
        out.put((t.numpy().tolist(), float(r.item()), float(tmax.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_counters_equal_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, rsum, tmax = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want, wsum = counters_for(range(GLOBAL_ENVS))
    assert got == want.tolist()
    assert abs(rsum - wsum) < 1e-9 * max(1.0, abs(wsum))
    assert tmax == 2.0
    assert got[11] == GLOBAL_ENVS * STEPS * len(ROUTES)
