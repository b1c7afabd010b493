"""The reference's own Python surface — its env.py over its pybind11 module MARLEnv (oracle/_ref, `make -C oracle pyref`)
— against the C-ABI driver around the same sources (RefEnv) that every parity test uses as its checker.  This pins the
checker to what a user of the reference actually calls, on the CPU."""
import numpy as np
import pytest

import pyoracle as po
import refpy_util as R

pytestmark = pytest.mark.skipif(not (R.have_pyref() and po.have_ref()), reason="oracle/_ref/MARLEnv.so not built (make -C oracle pyref)")

C1 = dict(num_agents=1, ego_routes=[("IN_6", "OUT_2")])
C2 = dict(num_agents=3, use_team_reward=True, ego_routes=[("IN_6", "OUT_2"), ("IN_4", "OUT_8"), ("IN_5", "OUT_7")])
C3 = dict(traffic_flow=True, traffic_density=0.5, ego_routes=[("IN_6", "OUT_2")])


def u32(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("cfg,steps,seed", [(C1, 2000, 0), (C2, 2000, 1), (C3, 1500, 2)], ids=["C1", "C2", "C3"])
def test_reference_env_py_over_marlenv_equals_c_driver(cfg, steps, seed):
    envmod = R.load_reference_env("MARLEnv")
    env = envmod.IntersectionEnv(dict(cfg))
    traffic = bool(cfg.get("traffic_flow", False))
    ref = po.RefEnv(3, cfg["ego_routes"], use_team=bool(cfg.get("use_team_reward", False)), traffic=traffic,
                    density=float(cfg.get("traffic_density", 0.5)), seed=seed, env_id=0)
    obs, info = env.reset()
    assert info == {} and (u32(obs) == u32(ref.obs()[0] if traffic else ref.obs())).all()
    rng = np.random.default_rng(seed)
    n = 1 if traffic else cfg["num_agents"]
    tick = 0
    for t in range(steps):
        a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        tick += 1
        R.marlenv_seed(seed, 0, tick)
        obs, rew, term, trunc, info = env.step(a[0] if traffic else a)
        r = ref.step(a)
        assert (u32(obs) == u32(r["obs"][0] if traffic else r["obs"])).all(), t
        assert (u32(rew) == u32(r["reward"][0] if traffic else r["reward"])).all(), t
        assert term == r["terminated"] and trunc == r["truncated"] and info["step"] == r["step"]
        assert info["status"] == [po.STATUS_NAMES[s] for s in r["status"]] and info["done"] == list(r["done"])
        assert info["agents_alive"] == r["agents_alive"]
        if traffic:
            assert len(env.traffic_cars) == len(ref.npcs())
        if term or trunc:
            env.reset()
            ref.reset()


def test_marlenv_state_surface():
    """What bindings.cpp:56-62,78-79,85-93 expose and the CUDA facade has to mirror: EnvState fields, the 72-beam lidars
    set_state leaves behind (IntersectionEnv.cpp:411-415), Car.update / check_collision."""
    envmod = R.load_reference_env("MARLEnv")
    env = envmod.IntersectionEnv(dict(C2))
    for _ in range(30):
        env.step(np.full((3, 2), 0.3, np.float32))
    core = env.env
    s = core.get_state()
    assert len(s.cars) == 3 and s.agent_ids == [1, 2, 3] and s.next_agent_id == 4 and s.step_count == 30
    assert [l.rays for l in core.lidars] == [96, 96, 96] and len(s.cars[0].path) == 160
    core.set_state(s)
    assert [l.rays for l in core.lidars] == [72, 72, 72]
    o = np.asarray(core.get_observations(), np.float32)
    assert (o[:, 31:103] == 1.0).all() and (o[:, 103:] == 0.0).all()
    M = R.marlenv_module()
    a, b = M.Car(), M.Car()
    st = M.State(); st.x, st.y, st.v, st.heading = 100.0, 100.0, 2.0, 0.5
    a.state = st
    st2 = M.State(); st2.x, st2.y, st2.heading = 130.0, 110.0, 2.0
    b.state = st2
    assert a.check_collision(b) is True
    a.update(0.5, -0.3, 1.0 / 60.0)
    u = po.ref_unit().car_update([100.0, 100.0, 2.0, 0.5, 0.0, 0.0], 0.5, -0.3, 1.0 / 60.0)
    assert (u32([a.state.x, a.state.y, a.state.v, a.state.heading]) == u32(u[:4])).all()
