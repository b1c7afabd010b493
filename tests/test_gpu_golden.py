"""CUDA stepper against the committed golden fixtures (made from the reference's own C++ by tests/golden/make_golden.py).
Needs no checker library at run time: this is the parity proof that survives on a box where neither /root/reference nor
oracle/_ref exists."""
import hashlib
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden as mg  # noqa: E402
import pyoracle as po  # noqa: E402  (only for the host copy of the Philox action stream)

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype == np.float32:
        return (a.view(np.uint32) == b.view(np.uint32)).all()
    return (a == b).all()


@pytest.mark.parametrize("name", list(mg.CASES))
def test_cuda_reproduces_reference_fixture(name):
    import torch
    from marl_traffic_intersection_b200 import BatchedIntersectionEnv
    kw = mg.CASES[name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    n = len(kw["ego_routes"])
    b = BatchedIntersectionEnv(dict(num_envs=1, num_agents=n, num_lanes=kw["num_lanes"], ego_routes=kw["ego_routes"],
                                    use_team_reward=kw.get("use_team", False), respawn_enabled=kw.get("respawn", True),
                                    max_steps=kw.get("max_steps", 2000), traffic_flow=kw.get("traffic", False),
                                    traffic_density=kw.get("density", 0.5), lidar_rays=kw.get("lidar_rays", 96),
                                    seed=mg.SEED, env_id_base=mg.ENV_ID, npc_capacity=32))
    h, lh = hashlib.sha256(), hashlib.sha256()
    obs, _ = b.reset()
    h.update(obs[0].cpu().numpy().tobytes())
    R = b.lidar_rays
    tick = 0
    for t in range(mg.STEPS):
        tick += 1
        a = po.philox_actions(mg.SEED, mg.ENV_ID, tick, n)
        o, rew, term, trunc, info = b.step(torch.from_numpy(a[None]).cuda())
        torch.cuda.synchronize()
        ob = o[0].cpu().numpy()
        h.update(ob.tobytes())
        assert _same(rew[0].cpu().numpy(), gold["reward"][t]), (name, t, "reward")
        assert _same(info["status"][0].cpu().numpy().astype(np.int8), gold["status"][t]), (name, t, "status")
        assert (int(term[0]), int(trunc[0]), int(info["agents_alive"][0])) == tuple(gold["flags"][t]), (name, t)
        ev = info["events"][0].cpu().numpy()
        got = (int(ev[0]), int(ev[1]), int(ev[2]), int(ev[3]) & 0xFFFFFFFF, int(ev[5]))
        assert got == tuple(int(x) for x in gold["events"][t]), (name, t, got, gold["events"][t])
        k = info["lidar_hit"][0, :, :R].cpu().numpy().astype(np.float32)
        lh.update(np.where(k == 0, np.float32(250.0), k * np.float32(4.0)).astype(np.float32).tobytes())
        if f"obs_{t + 1}" in gold.files:
            assert _same(ob, gold[f"obs_{t + 1}"]), (name, t, "obs")
        if bool(term[0]) or bool(trunc[0]):
            b.reset()
    assert (np.frombuffer(h.digest(), np.uint8) == gold["obs_sha256"]).all(), "sha256 over every obs of the rollout"
    assert (np.frombuffer(lh.digest(), np.uint8) == gold["lidar_sha256"]).all(), "sha256 over every lidar distance"
    fe = np.stack([b.buf[f][0].cpu().numpy() for f in ("ego_x", "ego_y", "ego_v", "ego_heading", "ego_steer", "ego_prev_dist")])
    assert _same(fe.astype(np.float32), gold["final_ego"])
    b.close()
