"""DRGBT's per-iteration geometry over the batched interface (rpmplv2_b200/drgbt.py).

The checker is the same `generate_gbur` driven by an oracle-backed space (computeDistance / extendGenSpine restated in
oracle/rvs_oracle.hpp), i.e. what the reference's DRGBT would compute one call at a time (DRGBT.cpp:101, 272-320, 475-495)."""
import numpy as np
import pytest

import scenes
from rpmplv2_b200 import drgbt, robots, synthetic


from oracle.oracle_space import OracleSpace


def _horizon(rb, q_current, H, seed):
    return synthetic.horizon_states(rb, q_current, H, seed)


def test_oracle_gbur_reached_states_are_collision_free_and_weights_follow_the_heuristic(ob):
    rb = robots.spatial(22)
    sc = scenes.dynamic_scene(rb, 5, 21, r_lo=0.45 * 2.2, r_hi=0.95 * 2.2, size=0.3 * 2.2)
    o = ob.Oracle(rb, robots.default_params())
    half = np.float32(0.5) * sc["dim"]
    o.set_obstacles(sc["pos"] - half, sc["pos"] + half, sc["is_ground"], np.full(6, 1000.0, np.float32))
    cand = synthetic.configs(rb, 300, seed=0xD6)
    q_cur = cand[int(np.argmax(o.distance(cand)[0]))]
    hor = _horizon(rb, q_cur, 52, 3)
    g = drgbt.generate_gbur(OracleSpace(o), q_cur, hor)
    assert g["d_c_current"] > 0 and g["q_reached"].shape == (52, 22)
    assert (g["d_c"] > 0).all()                                                       # the bur property: every reached state is free
    assert ((g["status"] == drgbt.STATUS_CRITICAL) == (g["d_c"] < 0.05)).all()
    moved = np.linalg.norm(g["q_reached"] - q_cur, axis=1)
    assert (moved > 0).all() and (moved <= np.linalg.norm(hor - q_cur, axis=1) + 1e-4).all()
    idx = np.where(np.arange(52) < 10, np.arange(52), -1)
    w, best, dmm = drgbt.horizon_weights(g["q_reached"], g["d_c"], g["d_c"], idx, g["status"], q_goal=hor[9], d_max_mean=0.0, num_iterations=0)
    assert w.shape == (52,) and (w >= 0).all() and (w <= 1).all() and dmm == pytest.approx(float(g["d_c"].max()))
    assert (w[idx == -1] <= 0.5 + 1e-6).all() and (w[g["status"] == drgbt.STATUS_CRITICAL] == 0).all()
    assert best == int(np.argmax(w))


@pytest.mark.gpu
@pytest.mark.parametrize("robot,H", [("spatial22", 52), ("spatial22", 220), ("xarm6", 20)])
def test_gpu_gbur_matches_the_oracle_backed_driver(ob, robot, H):
    from rpmplv2_b200.engine import Engine
    rb = robots.spatial(22) if robot == "spatial22" else robots.xarm6()
    reach = 2.2 if robot == "spatial22" else 1.0
    sc = scenes.dynamic_scene(rb, 5, 21, r_lo=0.45 * reach, r_hi=0.95 * reach, size=0.3 * reach)
    half = np.float32(0.5) * sc["dim"]
    mn, mx, tol = sc["pos"] - half, sc["pos"] + half, np.full(6, 1000.0, np.float32)
    o = ob.Oracle(rb, robots.default_params())
    o.set_obstacles(mn, mx, sc["is_ground"], tol)
    eng = Engine(rb)
    eng.set_obstacles(mn, mx, sc["is_ground"], tol)
    cand = synthetic.configs(rb, 300, seed=0xD6)
    d0 = o.distance(cand)[0]
    for k, q_cur in enumerate([cand[int(np.argmax(d0))], cand[int(np.argsort(d0)[-20])], cand[int(np.argmin(d0))]]):
        hor = _horizon(rb, q_cur, H, 7 + k)
        ge = drgbt.generate_gbur(eng, q_cur, hor)
        go = drgbt.generate_gbur(OracleSpace(o), q_cur, hor)
        assert ge["d_c_current"] == pytest.approx(go["d_c_current"], abs=1e-6)
        assert np.array_equal(ge["spine_status"], go["spine_status"]) and np.array_equal(ge["is_reached"], go["is_reached"])
        assert np.abs(ge["q_reached"] - go["q_reached"]).max() <= 1e-5
        assert np.abs(ge["d_c"] - go["d_c"]).max() <= 1e-4
        assert np.array_equal(ge["status"], go["status"]) or np.abs(go["d_c"] - 0.05).min() < 1e-4
    eng.close()


def test_horizon_weights_equal_a_literal_loop_restatement():
    """drgbt.horizon_weights (vectorised) against DRGBT::computeNextState (DRGBT.cpp:499-561) written out state by state."""
    rng = np.random.default_rng(12)
    H, n = 37, 6
    q_reached = rng.uniform(-2, 2, (H, n)).astype(np.float32)
    q_goal = rng.uniform(-2, 2, n).astype(np.float32)
    d_c = rng.uniform(0.0, 0.6, H).astype(np.float32)
    d_prev = (d_c + rng.normal(0, 0.05, H)).astype(np.float32)
    index = np.where(rng.random(H) < 0.4, -1, np.arange(H))
    status = np.where(d_c < 0.05, drgbt.STATUS_CRITICAL, drgbt.STATUS_GOOD)
    for num_iter, dmm in ((0, 0.0), (17, 0.31)):
        w, best, dmm_new = drgbt.horizon_weights(q_reached, d_c, d_prev, index, status, q_goal, dmm, num_iter)
        f = np.float32
        dist_to_goal = [f(np.sqrt(((q_reached[i] - q_goal) ** 2).sum(dtype=np.float32))) for i in range(H)]
        d_goal_min, d_c_max = min(dist_to_goal), max(d_c)
        weights_dist = [f(d_goal_min / dist_to_goal[i]) for i in range(H)]
        d_max_mean = f((f(num_iter) * f(dmm) + d_c_max) / f(num_iter + 1))
        mean = f(sum(float(x) for x in weights_dist) / H)
        max_weight, idx_best = 0.0, -1
        for i in range(H):
            if status[i] == drgbt.STATUS_CRITICAL:
                assert w[i] == 0
                continue
            weight = f(d_c[i] / d_max_mean + (d_c[i] - d_prev[i]) / d_max_mean + weights_dist[i] - mean)
            weight = min(max(weight, f(0)), f(1))
            if index[i] == -1:
                weight = f(weight * f(0.5))
            assert abs(float(w[i]) - float(weight)) <= 2e-6
            if weight > max_weight:
                max_weight, idx_best = weight, i
        assert best == idx_best and dmm_new == pytest.approx(float(d_max_mean))
