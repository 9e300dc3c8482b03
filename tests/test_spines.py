"""Row f-1: batched generalized-bur spines (RBTConnect::extendSpine + RGBTConnect::extendGenSpine, distance-to-planes variant).

CPU part: the oracle's restatement against (a) an independent numpy float32 re-expression of the recurrences in SURVEY.md §8.3
built only from the oracle's primitive outputs, (b) the property the burs of free C-space are built on — every state a spine
reaches is collision-free — and (c) the status rules. GPU part: rvs_extend_spines against the oracle.
Eigen's summation order inside R.col(i).dot(delta) is an assumption (Q24), so q_reached is compared with a tolerance.
"""
import numpy as np
import pytest

import scenes
from rpmplv2_b200 import engine, robots, synthetic

ADVANCED, TRAPPED, REACHED = 1, 2, 3


def _root_and_targets(ob, rb, mn, mx, g, tol, S, seed=11, spread=1.5):
    o = ob.Oracle(rb, robots.default_params())
    o.set_obstacles(mn, mx, g, tol)
    cand = synthetic.configs(rb, 400)
    d, p, near = o.distance(cand, robots.DIST_FULL)
    k = int(np.argmax(d))                                             # a root with good clearance
    rng = np.random.default_rng(seed)
    targets = (cand[k] + rng.normal(0, spread, (S, rb.num_dofs))).astype(np.float32)
    targets[0] = cand[k] + np.float32(1e-3)                           # inside the bubble: must be Reached in one step
    return o, cand[k].copy(), p[k].copy(), float(d[k]), near[k].copy(), targets


def _numpy_gen_spine(o, rb, q_root, prof_root, d_root, planes, q_e, sp):
    """SURVEY §8.3 in numpy float32 (control flow only; skeleton / radii / underestimation come from the oracle's primitives)."""
    f = np.float32
    n = rb.num_dofs
    if not d_root > 0:
        return TRAPPED, q_root.copy(), d_root
    q_cur, prof_cur, dc_cur, status, q_new, d = q_root.copy(), prof_root.copy(), f(d_root), 0, q_root.copy(), f(d_root)
    for _ in range(sp.num_layers):
        sk = o.skeleton(q_cur[None])[0]
        rho_prof, rho, q_temp, status, counter = np.zeros(n, f), f(0), q_cur.copy(), ADVANCED, 0
        while True:
            R = o.enclosing_radii(q_temp[None])[0]                    # n x (n+1)
            delta = np.abs(q_e - q_temp).astype(f)
            with np.errstate(divide="ignore", invalid="ignore"):
                if sp.use_expanded_bubble:
                    step = f(np.inf)
                    for i in range(n):
                        v = f(prof_cur[i] - rho_prof[i]) / f(np.dot(R[:, i + 1].astype(np.float64), delta.astype(np.float64)))
                        if v < step:
                            step = v
                else:
                    step = f(dc_cur - rho) / f(np.dot(R[:, n].astype(np.float64), delta.astype(np.float64)))
            if step > 1:
                q_new, status = q_e.copy(), REACHED
            else:
                q_new = (q_temp + step * (q_e - q_temp)).astype(f)
            counter += 1
            if counter == sp.num_iter_spine or status != ADVANCED:
                break
            sk_new = o.skeleton(q_new[None])[0]
            rho_k = np.linalg.norm((sk - sk_new).astype(np.float64), axis=1).astype(f)
            rho_prof = np.maximum(np.concatenate([[f(0)], rho_k[1:-1]]), rho_k[1:]).astype(f)
            rho = max(rho, rho_k[1:].max())
            q_temp = q_new
        du, pu, st = o.underestimation(q_new[None], planes[None], None)
        d = f(du[0])
        if d < sp.d_crit or status != ADVANCED:
            break
        q_cur, prof_cur, dc_cur = q_new, pu[0], d
    return status, q_new, d


@pytest.mark.parametrize("case,iters,expanded", [("xarm6", 1, 1), ("xarm6", 3, 1), ("xarm6", 2, 0), ("planar2", 1, 1), ("spatial10", 2, 1)])
def test_oracle_spines_follow_the_recurrences_and_stay_collision_free(ob, case, iters, expanded):
    if case == "xarm6":
        rb, mn, mx, g, tol = scenes.xarm6_scene(51, "D2", 0.45)
    elif case == "planar2":
        rb, mn, mx, g, tol = scenes.planar2_scenario_test()
    else:
        rb, mn, mx, g, tol = scenes.chain_scene(robots.spatial(10), 6)
    tol = np.full_like(tol, 1000.0)                                   # planes of every obstacle (as the dynamic planners configure it)
    S = 60
    o, q_root, prof, d_root, planes, targets = _root_and_targets(ob, rb, mn, mx, g, tol, S)
    sp = robots.RvsSpineParams(5, iters, expanded, 0.05)
    q_out, d_out, status = o.extend_spines(q_root, prof, d_root, planes, targets, sp)
    assert status[0] == REACHED and (q_out[0] == targets[0]).all()
    assert set(np.unique(status)) <= {ADVANCED, REACHED} and (status == ADVANCED).sum() > 5
    # the bur property: every reached state is collision-free in the REAL scene
    d_real, _, _ = o.distance(q_out, robots.DIST_MIN)
    assert (d_real > 0).all()
    # distance-to-planes never overestimates the real distance (up to float noise)
    assert (d_out <= d_real + 1e-4).all()
    # every reached state lies on the segment root -> target
    t = np.einsum("sd,sd->s", q_out - q_root, targets - q_root) / np.maximum(np.einsum("sd,sd->s", targets - q_root, targets - q_root), 1e-12)
    assert (t > -1e-5).all() and (t < 1 + 1e-5).all()
    assert np.abs((q_root + t[:, None] * (targets - q_root)) - q_out).max() < 1e-4
    for s in range(0, S, 3):
        st, qn, dn = _numpy_gen_spine(o, rb, q_root, prof, d_root, planes, targets[s], sp)
        assert st == status[s]
        assert np.abs(qn - q_out[s]).max() < 2e-5 and abs(float(dn) - float(d_out[s])) < 2e-5
    # a root without clearance is Trapped and returned unchanged (RBTConnect.cpp:91-92)
    q2, d2, s2 = o.extend_spines(q_root, prof, 0.0, planes, targets[:3], sp)
    assert (s2 == TRAPPED).all() and (q2 == q_root).all() and (d2 == 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("case,iters,expanded,layers", [("xarm6", 1, 1, 5), ("xarm6", 3, 1, 5), ("xarm6", 2, 0, 3), ("planar2", 1, 1, 5),
                                                        ("spatial22", 1, 1, 5), ("spatial22", 2, 0, 2)])
def test_gpu_spines_vs_oracle(ob, case, iters, expanded, layers):
    if case == "xarm6":
        rb, mn, mx, g, tol = scenes.xarm6_scene(51, "D2", 0.45)
    elif case == "planar2":
        rb, mn, mx, g, tol = scenes.planar2_scenario_test()
    else:
        rb, mn, mx, g, tol = scenes.chain_scene(robots.spatial(22), 6)
    tol = np.full_like(tol, 1000.0)
    S = 262                                                           # the largest spine set a planner builds
    o, q_root, prof, d_root, planes, targets = _root_and_targets(ob, rb, mn, mx, g, tol, S)
    sp = robots.RvsSpineParams(layers, iters, expanded, 0.05)
    eng = engine.Engine(rb)
    eng.set_obstacles(mn, mx, g, tol)
    # the root quantities come from the engine itself, as a planner would obtain them
    d_e, p_e, near_e = eng.distance(q_root[None], robots.DIST_FULL)
    assert abs(float(d_e[0]) - d_root) <= 1e-4
    q_g, d_g, s_g = eng.extend_spines(q_root, p_e[0], float(d_e[0]), near_e[0], targets, sp)
    q_o, d_o, s_o = o.extend_spines(q_root, p_e[0], float(d_e[0]), near_e[0], targets, sp)
    assert (s_g == s_o).all()
    assert np.abs(q_g - q_o).max() <= 1e-5 and np.abs(d_g - d_o).max() <= 1e-4
    assert (q_g == q_o).mean() > 0.99                                 # in practice bit-identical
    assert (eng.distance(q_g, robots.DIST_MIN)[0] > 0).all()          # collision-free in the real scene
    q_t, d_t, s_t = eng.extend_spines(q_root, p_e[0], 0.0, near_e[0], targets[:5], sp)
    assert (s_t == TRAPPED).all() and (q_t == q_root).all()
