"""Drop-in proof: the reference's OWN env.py, unmodified (byte-compiled under oracle/_ref/refpy by
`make -C oracle pyref`), is executed twice — once over the reference's pybind11 module MARLEnv, once with
`sys.modules["cpp_backend"]` = this repo's cpp_backend (the CUDA stepper behind the C ABI) — and everything reset()/step()
return, the info dict, get_state()/set_state() and the lidars must agree bit for bit."""
import numpy as np
import pytest

import pyoracle as po
import refpy_util as R

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not R.have_pyref(), reason="oracle/_ref/MARLEnv.so not built (make -C oracle pyref)")]

C1 = dict(num_agents=1, ego_routes=[("IN_6", "OUT_2")])
C2 = dict(num_agents=3, use_team_reward=True, ego_routes=[("IN_6", "OUT_2"), ("IN_4", "OUT_8"), ("IN_5", "OUT_7")])
C3 = dict(traffic_flow=True, traffic_density=0.5, ego_routes=[("IN_6", "OUT_2")])
C2B = dict(num_agents=4, num_lanes=2, respawn_enabled=False, max_steps=150,
           reward_config=dict(progress_scale=3.0, crash_vehicle_penalty=-7.0, team_alpha=0.6), use_team_reward=True)


def u32(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def pair(cfg):
    from marl_traffic_intersection_b200 import cpp_backend
    ref_mod = R.load_reference_env("MARLEnv")
    cuda_mod = R.load_reference_env(cpp_backend)
    assert ref_mod.cpp_backend is not cuda_mod.cpp_backend and cuda_mod.cpp_backend is cpp_backend
    return ref_mod.IntersectionEnv(dict(cfg)), cuda_mod.IntersectionEnv(dict(cfg))


def same_step(out_r, out_c, where):
    (o1, r1, t1, tr1, i1), (o2, r2, t2, tr2, i2) = out_r, out_c
    assert type(o1) is type(o2) and o1.shape == o2.shape and o1.dtype == o2.dtype, where
    assert (u32(o1) == u32(o2)).all(), where
    assert type(r1) is type(r2) and (u32(r1) == u32(r2)).all(), where
    assert (t1, tr1) == (t2, tr2) and type(t2) is bool and type(tr2) is bool, where
    assert list(i1) == list(i2), where                       # same keys, same order
    for k in i1:
        if k == "rewards":
            assert type(i1[k]) is type(i2[k]) and (u32(i1[k]) == u32(i2[k])).all(), (where, k)
        else:
            assert i1[k] == i2[k] and type(i1[k]) is type(i2[k]), (where, k)


@pytest.mark.parametrize("cfg,steps", [(C1, 700), (C2, 700), (C3, 900), (C2B, 500)], ids=["C1", "C2", "C3_traffic", "lanes2_norespawn"])
def test_reference_env_py_runs_unchanged_on_the_cuda_backend(cfg, steps):
    ref, cud = pair(cfg)
    traffic = bool(cfg.get("traffic_flow", False))
    n = 1 if traffic else cfg["num_agents"]
    (o1, i1), (o2, i2) = ref.reset(), cud.reset()
    assert (u32(o1) == u32(o2)).all() and i1 == i2 == {}
    assert len(ref.cars) == len(cud.cars) == n and ref.points == cud.points
    rng = np.random.default_rng(7)
    tick = 0
    for t in range(steps):
        a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        if t % 97 == 5:
            a[:] = 0.0                                        # throttle == 0 branch of Car::update
        act = a[0] if traffic else a
        tick += 1
        R.marlenv_seed(0, 0, tick)                            # the CUDA facade's stream: seed 0, env 0, tick = steps so far
        out_r, out_c = ref.step(act), cud.step(act)
        same_step(out_r, out_c, t)
        if traffic:
            assert len(ref.traffic_cars) == len(cud.traffic_cars)
            for c1, c2 in zip(ref.traffic_cars, cud.traffic_cars):
                assert (u32([c1.state.x, c1.state.y, c1.state.v, c1.state.heading]) == u32([c2.state.x, c2.state.y, c2.state.v, c2.state.heading])).all()
                assert c1.intention == c2.intention and c1.path_index == c2.path_index
        if t % 50 == 0:                                       # by-value views of bindings.cpp:60-62
            for c1, c2 in zip(ref.env.cars, cud.env.cars):
                assert (u32([c1.state.x, c1.state.y, c1.state.v, c1.state.heading]) == u32([c2.state.x, c2.state.y, c2.state.v, c2.state.heading])).all()
                assert (c1.alive, c1.intention, c1.path_index, c1.length, c1.width) == (c2.alive, c2.intention, c2.path_index, c2.length, c2.width)
                assert (u32(np.asarray(c1.path)) == u32(np.asarray(c2.path))).all()
            for l1, l2 in zip(ref.env.lidars, cud.env.lidars):
                assert l1.rays == l2.rays and l1.distances == l2.distances and (u32(l1.rel_angles) == u32(l2.rel_angles)).all()
                assert (u32(l1.normalized()) == u32(l2.normalized())).all()
            assert ref.env.step_count == cud.env.step_count
        if out_r[2] or out_r[3]:
            (o1, _), (o2, _) = ref.reset(), cud.reset()
            assert (u32(o1) == u32(o2)).all(), ("reset", t)
    assert ref.render() is None and cud.render() is None      # render_mode is None: env.py:210-212
    cud.close()


@pytest.mark.parametrize("cfg", [C2, C3], ids=["C2", "C3_traffic"])
def test_get_state_set_state_rollback(cfg):
    """The reference's MCTS-rollback use (EnvState.h:3-15): snapshot, explore, restore, replay — through the Python
    objects of bindings.cpp:56-62,78-79, including the default 72-beam lidars set_state leaves behind (:411-415)."""
    ref, cud = pair(cfg)
    traffic = bool(cfg.get("traffic_flow", False))
    n = 1 if traffic else cfg["num_agents"]
    rng = np.random.default_rng(11)
    acts = rng.uniform(-1, 1, (400, n, 2)).astype(np.float32)
    tick = 0

    def both(t):
        nonlocal tick
        act = acts[t][0] if traffic else acts[t]
        tick += 1
        R.marlenv_seed(0, 0, tick)
        o = ref.step(act), cud.step(act)
        same_step(o[0], o[1], t)
        return o[0]
    for t in range(120):
        both(t)
    s_ref, s_cud = ref.env.get_state(), cud.env.get_state()
    assert (s_ref.step_count, s_ref.next_agent_id, list(s_ref.agent_ids)) == (s_cud.step_count, s_cud.next_agent_id, list(s_cud.agent_ids))
    assert len(s_ref.cars) == len(s_cud.cars) and len(s_ref.traffic_cars) == len(s_cud.traffic_cars)
    for t in range(120, 160):                                 # explore
        both(t)
    ref.env.set_state(s_ref)
    cud.env.set_state(s_cud)
    assert [l.rays for l in cud.env.lidars] == [l.rays for l in ref.env.lidars] == [72] * n
    assert (u32(np.asarray(ref.env.get_observations())) == u32(np.asarray(cud.env.get_observations()))).all()
    assert ref.env.step_count == cud.env.step_count == 120
    for t in range(160, 330):                                 # replay from the restored state (now with 72 beams)
        out = both(t)
        if out[2] or out[3]:
            break
    # an edited snapshot: move ego 0, kill nobody, roll the step counter
    s_ref, s_cud = ref.env.get_state(), cud.env.get_state()
    for s in (s_ref, s_cud):
        c = s.cars[0]
        st = c.state
        st.x, st.y, st.v = 375.0, 600.0, 1.5
        c.state = st
        s.cars = [c] + list(s.cars[1:])
        s.step_count = 7
    ref.env.set_state(s_ref)
    cud.env.set_state(s_cud)
    for t in range(330, 400):
        both(t)
    (o1, _), (o2, _) = ref.reset(), cud.reset()               # reset() + add_car_with_route: back to 96 beams
    assert (u32(o1) == u32(o2)).all() and [l.rays for l in cud.env.lidars] == [l.rays for l in ref.env.lidars] == [96] * n
    cud.close()


def test_car_unit_methods_and_live_reconfiguration():
    from marl_traffic_intersection_b200 import cpp_backend
    M = R.marlenv_module()
    rng = np.random.default_rng(3)
    for _ in range(40):
        x, y, h, v = rng.uniform(100, 600), rng.uniform(100, 600), rng.uniform(-3.1, 3.1), rng.uniform(0, 8)
        thr, st, dt = (0.0 if rng.random() < 0.2 else rng.uniform(-1, 1)), rng.uniform(-2, 2), 1.0 / 60.0
        a, b = M.Car(), cpp_backend.Car()
        s1 = M.State(); s1.x, s1.y, s1.v, s1.heading = x, y, v, h
        a.state = s1
        b.state = cpp_backend.State(np.float32(x), np.float32(y), np.float32(v), np.float32(h))
        for _k in range(3):
            a.update(thr, st, dt)
            b.update(thr, st, dt)
        assert (u32([a.state.x, a.state.y, a.state.v, a.state.heading]) == u32([b.state.x, b.state.y, b.state.v, b.state.heading])).all()
        o1, o2 = M.Car(), cpp_backend.Car()
        s2 = M.State(); s2.x, s2.y, s2.heading = a.state.x + rng.uniform(-60, 60), a.state.y + rng.uniform(-60, 60), rng.uniform(-3, 3)
        o1.state = s2
        o2.state = cpp_backend.State(s2.x, s2.y, 0.0, s2.heading)
        assert a.check_collision(o1) == b.check_collision(o2)
    # reward weights / configure() changed on the live env object take effect without rebuilding the device handle
    ref, cud = pair(C2)
    handle = cud.env._benv
    for e in (ref, cud):
        e.env.reward_config.k_prog = 2.5
        e.env.reward_config.alpha = 0.5
        e.env.configure(True, False, 40)
    for t in range(60):
        act = rng.uniform(-1, 1, (3, 2)).astype(np.float32)
        o = ref.step(act), cud.step(act)
        same_step(o[0], o[1], t)
        if o[0][2] or o[0][3]:
            ref.reset(); cud.reset()
    assert cud.env._benv is handle
    cud.close()
