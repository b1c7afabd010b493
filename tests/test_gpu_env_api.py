"""The reference's Python surface (env.py / cpp_backend.py) on top of the CUDA stepper: shapes, dtypes, info keys,
exception types — the drop-in contract of SURVEY.md §8b."""
import numpy as np
import pytest

import pyoracle as po

pytestmark = pytest.mark.gpu


def Checker(*a, **k):
    """The reference's own C++ when oracle/_ref travelled with the repo, else the pinned C port."""
    return (po.RefEnv if po.have_ref() else po.OracleEnv)(*a, **k)


def test_multi_agent_surface_and_values():
    from marl_traffic_intersection_b200 import IntersectionEnv
    routes = [("IN_6", "OUT_2"), ("IN_4", "OUT_8"), ("IN_5", "OUT_7")]
    env = IntersectionEnv({"num_agents": 3, "use_team_reward": True, "ego_routes": routes})
    ref = Checker(3, routes, use_team=True)
    obs, info = env.reset()
    assert obs.shape == (3, 127) and obs.dtype == np.float32 and info == {}
    assert (obs.view(np.uint32) == ref.obs().view(np.uint32)).all()
    rng = np.random.default_rng(0)
    for _ in range(150):
        a = rng.uniform(-1, 1, (3, 2)).astype(np.float32)
        obs, rew, term, trunc, info = env.step(a)
        r = ref.step(a)
        assert obs.shape == (3, 127) and rew.shape == (3,) and rew.dtype == np.float32
        assert isinstance(term, bool) and isinstance(trunc, bool)
        assert set(info) == {"step", "rewards", "collisions", "agents_alive", "terminated", "truncated", "done", "status"}
        assert (obs.view(np.uint32) == r["obs"].view(np.uint32)).all() and (rew.view(np.uint32) == r["reward"].view(np.uint32)).all()
        assert info["status"] == [po.STATUS_NAMES[s] for s in r["status"]] and info["done"] == list(r["done"])
        assert info["collisions"] == {i + 1: po.STATUS_NAMES[s] for i, s in enumerate(r["status"])}
        assert info["step"] == r["step"] and info["agents_alive"] == r["agents_alive"]
    assert len(env.cars) == 3 and env.cars[0].state.x == 720.0      # reset-time snapshot, as in the reference (env.py:152)
    assert env.env.cars[0].state.x == float(ref.egos()[0]["x"])   # a fresh by-value copy, like the pybind property
    with pytest.raises(ValueError):
        env.step(np.zeros(6, np.float32))          # env.py:178
    env.close()


def test_traffic_mode_forces_single_agent_and_scalar_returns():
    from marl_traffic_intersection_b200 import IntersectionEnv
    env = IntersectionEnv({"traffic_flow": True, "num_agents": 5, "traffic_density": 2.0, "seed": 3})
    assert env.num_agents == 1                          # env.py:87-90
    obs, _ = env.reset()
    assert obs.shape == (127,)
    ref = Checker(3, [("IN_1", "OUT_4")], traffic=True, density=2.0, seed=3)
    for _ in range(200):
        obs, rew, term, trunc, info = env.step([0.3, 0.0])
        r = ref.step([[0.3, 0.0]])
        assert obs.shape == (127,) and isinstance(rew, float) and isinstance(info["rewards"], float)
        assert (obs.view(np.uint32) == r["obs"][0].view(np.uint32)).all()
        assert len(env.traffic_cars) == len(ref.npcs())
    env.close()


def test_error_behaviour():
    from marl_traffic_intersection_b200 import BatchedIntersectionEnv, IntersectionEnv, cpp_backend
    with pytest.raises(IndexError):
        IntersectionEnv({"num_agents": 1, "ego_routes": [("IN_6", "OUT_99")]})        # RouteGen.cpp:120 -> IndexError
    e = cpp_backend.IntersectionEnv(3)
    e.reset()
    e.add_car_with_route("IN_99", "OUT_2")                                           # silent no-op, IntersectionEnv.cpp:79-82
    e.add_car_with_route("IN_6", "OUT_2")
    assert len(e.get_observations()) == 1
    r = e.step([], [])                                                               # missing actions -> 0, :153-154
    assert len(r.obs) == 1 and r.agent_ids == [1] and r.step == 1
    with pytest.raises(ValueError):
        BatchedIntersectionEnv({"num_envs": 2, "num_agents": 2, "ego_routes": [("IN_6", "OUT_2")]})
    assert cpp_backend.has_cpp_backend()


R3 = po.ROUTES_3LANES


def _benv():
    from marl_traffic_intersection_b200 import BatchedIntersectionEnv
    return BatchedIntersectionEnv


def test_headless_render_shows_road_cars_and_hits():
    """isx_render: pixel classes at known places — road/grass exactly as the checker's road map away
    from cars and beams, every alive ego's centre in its palette colour, every lidar hit pixel red."""
    import torch
    cfg = dict(num_envs=3, num_agents=4, num_lanes=3, ego_routes=R3[:4], traffic_flow=True, traffic_density=4.0, seed=71, auto_reset=True)
    b = _benv()(cfg)
    b.rollout(140)
    img = b.render(1).cpu().numpy()
    assert img.shape == (750, 750, 3) and img.dtype == np.uint8
    road = po.ref_unit().road_map(3) if po.have_ref() else po.oracle_unit().road_map(3)
    road = np.asarray(road).reshape(750, 750).astype(bool)
    grass, surf = np.array([58, 125, 68]), np.array([70, 70, 74])
    is_grass, is_surf = (img == grass).all(-1), (img == surf).all(-1)
    assert not (is_grass & road).any() and not (is_surf & ~road).any()          # never the wrong base colour
    assert is_grass[~road].mean() > 0.95 and is_surf[road].mean() > 0.6
    pal = np.array([[231, 76, 60], [52, 152, 219], [46, 204, 113], [155, 89, 182]])
    ex, ey, al = (b.buf[k][1].cpu().numpy() for k in ("ego_x", "ego_y", "ego_alive"))
    hits = b.buf["lidar_hit"][1].cpu().numpy()
    seen_hit = 0
    for a in range(4):
        if not al[a]:
            continue
        cx, cy = int(ex[a]), int(ey[a])
        if 0 <= cx < 750 and 0 <= cy < 750:
            px = img[cy, cx]
            assert (px == pal[a]).all() or (px == [220, 30, 30]).all() or any((px == pal[o]).all() for o in range(4)) or (px == [128, 128, 128]).all() or (px == [20, 20, 20]).all() or (px == [250, 250, 250]).all(), (a, px)
        seen_hit += int((hits[a][:96] > 0).sum())
    assert seen_hit > 0 and ((img == [220, 30, 30]).all(-1)).sum() >= 5
    # the single-env facade returns the same kind of picture
    from marl_traffic_intersection_b200 import IntersectionEnv
    env = IntersectionEnv({"num_agents": 2, "ego_routes": R3[:2], "render_mode": "human"})
    pic = env.render()
    assert pic.shape == (750, 750, 3) and (pic[int(env.env.cars[0].state.y), int(env.env.cars[0].state.x)] == [231, 76, 60]).all()
    quiet = IntersectionEnv({"num_agents": 2, "ego_routes": R3[:2]})
    assert quiet.render() is None                      # env.py:210-212: nothing unless render_mode == "human"
    quiet.close()
    env.close()
    b.close()
