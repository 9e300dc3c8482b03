"""Row f-1 / f-4: RRT-Connect, RBT-Connect and RGBT-Connect on the batched state space (rpmplv2_b200/cpp/planning).

CPU part (oracle backend): the planners solve the reference's shipped scenarios (planar_2dof scenario_test, xarm6 scenario1,
spatial_10dof scenario1 — geometry retyped in tests/scenes.py), every returned path starts/ends at q_start/q_goal and is
collision-free when densely re-sampled, and running P runs in lockstep (one pool) gives exactly the runs a pool of one gives
(batching queries across runs must not change any run).
GPU part: the same planner code on the CUDA backend must reproduce the oracle-backed runs seed by seed (the two arms differ
only in who answers the state-space queries), and its paths must be collision-free according to the ORACLE.
"""
import numpy as np
import pytest

import scenes
from rpmplv2_b200 import engine, planning, robots

SCENARIOS = {
    # q_start / q_goal: data/planar_2dof/scenario_test/scenario_test.yaml:24-25, data/xarm6/scenario1/scenario1.yaml:31-32,
    # data/spatial_10dof/scenario1/scenario1.yaml:31-32 (of the reference tree)
    "planar2": (scenes.planar2_scenario_test, [-2, -2.5], [2, 2.5]),
    "xarm6": (scenes.xarm6_scenario1, [0, 0.7854, -2.3562, 0, 0, 0], [3.14159, 0.7854, -2.3562, 3.14159, 0, 3.14159]),
    "spatial10": (scenes.spatial10_scenario1, [0, 1.57, 0, 0, 0, 0, 0, 0, 0, 0], [3.14, 1.57, 0, 0, 0, 0, 0, 0, 0, 0]),
}


def _oracle(ob, name):
    fn, qs, qg = SCENARIOS[name]
    rb, mn, mx, g, tol = fn()
    o = ob.Oracle(rb, robots.default_params())
    o.set_obstacles(mn, mx, g, tol)
    return o, (rb, mn, mx, g, tol), np.asarray(qs, np.float32), np.asarray(qg, np.float32)


def _dense_path_is_free(o, path, step=0.01):
    pts = []
    for a, b in zip(path[:-1], path[1:]):
        k = max(2, int(np.ceil(np.linalg.norm(b - a) / step)) + 1)
        pts.append(a + (b - a) * np.linspace(0, 1, k, dtype=np.float32)[:, None])
    return bool(o.is_valid(np.concatenate(pts)).all())


@pytest.mark.parametrize("name", ["planar2", "xarm6", "spatial10"])
@pytest.mark.parametrize("kind", [planning.RRT_CONNECT, planning.RBT_CONNECT, planning.RGBT_CONNECT])
def test_oracle_backed_planners_solve_the_shipped_scenarios(ob, name, kind):
    o, _, qs, qg = _oracle(ob, name)
    cfg = planning.default_config(kind, max_planning_time=60.0, max_num_iter=20000)
    P = 4
    res, _ = ob.plan_solve_batch(o, cfg, qs, qg, seeds=np.arange(P) + 100, max_path=2048)
    assert all(r["success"] for r in res)
    for r in res:
        path = r["path"]
        assert r["path_len"] == len(path) >= 2 and r["states"] >= r["path_len"]
        assert (path[0] == qs).all() and (path[-1] == qg).all()
        assert _dense_path_is_free(o, path)
        if kind == planning.RRT_CONNECT:
            assert r["num_distances"] == 0 and r["num_spines"] == 0
            assert np.linalg.norm(np.diff(path, axis=0), axis=1).max() <= cfg.eps_step + 1e-4     # every edge is one EPS_STEP
        else:
            assert r["num_distances"] > 0 and r["num_spines"] > 0
    # lockstep over the P runs == P pools of one run
    res_l, st = ob.plan_solve_batch(o, cfg, qs, qg, seeds=np.arange(P) + 100, max_path=2048, lockstep=True)
    assert st["rounds"] > 0 and st["backend_calls"] >= st["rounds"]
    for a, b in zip(res, res_l):
        assert a["iterations"] == b["iterations"] and a["states"] == b["states"] and (a["path"] == b["path"]).all()
    # different seeds give different trees
    assert len({r["states"] for r in res}) > 1


def test_invalid_start_is_reported_not_planned(ob):
    o, _, qs, qg = _oracle(ob, "planar2")
    bad = np.array([np.pi / 4, 0], np.float32)            # link 2 reaches (1.41, 1.41)... link 1 passes through the box at (1, 1)
    assert not o.is_valid(bad[None])[0]
    res, _ = ob.plan_solve_batch(o, planning.default_config(planning.RGBT_CONNECT), np.stack([bad, qs]), np.stack([qg, qg]), seeds=[1, 2])
    assert res[0]["invalid_ends"] == 1 and not res[0]["success"] and res[0]["iterations"] == 0
    assert res[1]["success"] and res[1]["invalid_ends"] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["planar2", "xarm6", "spatial10"])
@pytest.mark.parametrize("kind", [planning.RRT_CONNECT, planning.RBT_CONNECT, planning.RGBT_CONNECT])
def test_gpu_backed_planners_reproduce_the_oracle_backed_runs(ob, name, kind):
    o, (rb, mn, mx, g, tol), qs, qg = _oracle(ob, name)
    eng = engine.Engine(rb)
    eng.set_obstacles(mn, mx, g, tol)
    cfg = planning.default_config(kind, max_planning_time=60.0, max_num_iter=20000)
    P = 12
    seeds = np.arange(P) + 7
    res_g, st = planning.solve_batch(eng, cfg, qs, qg, seeds=seeds, max_path=2048)
    res_o, _ = ob.plan_solve_batch(o, cfg, qs, qg, seeds=seeds, max_path=2048, threads=4)
    assert st["kernel_launches"] >= st["backend_calls"] > 0
    same = 0
    for a, b in zip(res_g, res_o):
        assert a["success"] and b["success"]
        assert (a["path"][0] == qs).all() and (a["path"][-1] == qg).all()
        assert _dense_path_is_free(o, a["path"])           # GPU-planned path checked by the CPU oracle
        if a["iterations"] == b["iterations"] and a["states"] == b["states"] and a["path_len"] == b["path_len"]:
            assert np.abs(a["path"] - b["path"]).max() <= 1e-4
            same += 1
    # the arms differ only by last-ulp effects of FK (device vs host sin/cos) on a few queries; a diverged run is still a valid run
    assert same >= P - 2
    eng.close()


@pytest.mark.gpu
def test_gpu_planner_pool_scales_in_rounds_not_runs():
    """64 runs in lockstep need about as many backend rounds as the longest single run, not 64 times as many."""
    fn, qs, qg = SCENARIOS["xarm6"]
    rb, mn, mx, g, tol = fn()
    eng = engine.Engine(rb)
    eng.set_obstacles(mn, mx, g, tol)
    cfg = planning.default_config(planning.RGBT_CONNECT, max_planning_time=60.0)
    res1, st1 = planning.solve_batch(eng, cfg, qs, qg, seeds=[3])
    res64, st64 = planning.solve_batch(eng, cfg, qs, qg, seeds=np.arange(64))
    assert all(r["success"] for r in res64)
    longest = max(r["rounds"] for r in res64)
    # runs drift apart (one asks for spines while another asks for a distance), so a round can need one call per query kind
    assert st64["backend_calls"] <= 2.5 * longest + 8
    assert st64["backend_calls"] < 0.1 * sum(r["rounds"] for r in res64)
    assert st1["backend_calls"] == res1[0]["rounds"]
    eng.close()


# --------------------------------------------------------------------------------------------------------------------------------
# Row f-3 pinned by the reference's own planners: RRTConnect / RBTConnect / RGBTConnect::solve() compiled unmodified on top of the
# reference's own Tree and RealVectorSpace (oracle/rvs_ref_capi.cpp: rvsref_plan; tests/golden/make_plan_golden.py ->
# tests/golden/plan_golden.npz).  With RVS_PLAN_RNG=libc a single run of rpmplv2_b200/cpp/planning draws from the same std::rand()
# stream as Eigen::VectorXf::Random does in the work-alike, and must then reproduce the reference run: success, iterations, number of
# states, and the path bit for bit.
# --------------------------------------------------------------------------------------------------------------------------------
def _plan_golden():
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("make_plan_golden", os.path.join(ROOT, "tests", "golden", "make_plan_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, dict(np.load(os.path.join(ROOT, "tests", "golden", "plan_golden.npz")))


def _same_run(r, counts, path, exact=True):
    if not (r["success"] == bool(counts[0]) and r["iterations"] == counts[1] and r["states"] == counts[2] and r["path_len"] == len(path)):
        return False
    return np.array_equal(r["path"], path) if exact else bool(np.abs(r["path"] - path).max() <= 1e-4)


def test_oracle_backed_planners_reproduce_the_reference_planner_runs(ob, monkeypatch):
    mod, gold = _plan_golden()
    monkeypatch.setenv("RVS_PLAN_RNG", "libc")
    for name in mod.SCENARIOS:
        o, _, qs, qg = _oracle(ob, name)
        for kind in mod.KINDS:
            for seed in mod.SEEDS:
                res, _ = ob.plan_solve_batch(o, mod.config(kind), qs, qg, seeds=[seed], max_path=mod.MAX_PATH)
                counts, path = gold[f"{name}_k{kind}_s{seed}_counts"], gold[f"{name}_k{kind}_s{seed}_path"]
                assert counts[0] == 1 and counts[3] + counts[4] == counts[2]
                assert _same_run(res[0], counts, path), (name, kind, seed, res[0]["iterations"], counts)
    # without the switch the planners use their own counter-based stream: a different (still valid) run
    monkeypatch.delenv("RVS_PLAN_RNG")
    o, _, qs, qg = _oracle(ob, "xarm6")
    res, _ = ob.plan_solve_batch(o, mod.config(planning.RGBT_CONNECT), qs, qg, seeds=[1], max_path=mod.MAX_PATH)
    assert res[0]["success"] and not _same_run(res[0], gold["xarm6_k2_s1_counts"], gold["xarm6_k2_s1_path"])


@pytest.mark.parametrize("chunk", [1, 4])
def test_connect_chains_do_not_change_a_run(ob, monkeypatch, chunk):
    """Request::CONNECT (RVS_PLAN_CONNECT_CHUNK steps of connectSpine / connectGenSpine's loop per query) against the golden runs of the
    reference's own planners: same iterations, states and paths as with one query per step."""
    mod, gold = _plan_golden()
    monkeypatch.setenv("RVS_PLAN_RNG", "libc")
    monkeypatch.setenv("RVS_PLAN_CONNECT_CHUNK", str(chunk))
    for name in mod.SCENARIOS:
        o, _, qs, qg = _oracle(ob, name)
        for kind in (planning.RBT_CONNECT, planning.RGBT_CONNECT):
            res, _ = ob.plan_solve_batch(o, mod.config(kind), qs, qg, seeds=[2], max_path=mod.MAX_PATH)
            assert _same_run(res[0], gold[f"{name}_k{kind}_s2_counts"], gold[f"{name}_k{kind}_s2_path"]), (name, kind, chunk)


def test_oracle_backed_planners_reproduce_the_compiled_reference_on_fresh_runs(ob, monkeypatch):
    """Same comparison against oracle/_ref/librvs_ref.so itself: other seeds and non-default parameters (plain bubble, 3 spine
    iterations, 3 layers, a coarser EPS_STEP), and the invalid-start rule."""
    import ctypes
    import os
    from conftest import ROOT
    so = os.path.join(ROOT, "oracle", "_ref", "librvs_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref is built where /root/reference exists")
    lib = ctypes.CDLL(so)
    if not hasattr(lib, "rvsref_plan"):
        pytest.skip("oracle/_ref/librvs_ref.so predates the planner harness")
    mod, _ = _plan_golden()
    monkeypatch.setenv("RVS_PLAN_RNG", "libc")
    variants = [dict(), dict(eps_step=0.25, max_extension_steps=10, num_spines=3, delta=1.5,
                             spine=robots.RvsSpineParams(3, 3, 0, 0.03))]
    for name in mod.SCENARIOS:
        h, rb, qs, qg = mod.open_reference(lib, name)
        o, _, _, _ = _oracle(ob, name)
        for kind in mod.KINDS:
            for vi, over in enumerate(variants):
                cfg = mod.config(kind, **over)
                for seed in (11 + vi, 29 + vi):
                    want = mod.run_reference(lib, h, rb.num_dofs, cfg, qs, qg, seed)
                    res, _ = ob.plan_solve_batch(o, cfg, qs, qg, seeds=[seed], max_path=mod.MAX_PATH)
                    counts = (want["success"], want["iterations"], want["states"])
                    assert _same_run(res[0], counts, want["path"]), (name, kind, vi, seed, res[0]["iterations"], counts)
        if name == "xarm6":
            bad = np.array([0, 1.5, 0, 0, 0, 0], np.float32)                     # arm through the table
            assert mod.run_reference(lib, h, rb.num_dofs, mod.config(planning.RRT_CONNECT), bad, qg, 1)["invalid_ends"]
            res, _ = ob.plan_solve_batch(o, mod.config(planning.RRT_CONNECT), bad, qg, seeds=[1])
            assert res[0]["invalid_ends"] and not res[0]["success"]
        lib.rvsref_destroy(h)


def _same_log_states(ours, theirs):
    if len(ours) != len(theirs):
        return False
    for (q, par), (q_r, par_r) in zip(ours, theirs):
        if (par is None) != (par_r is None) or not np.allclose(q, q_r, rtol=2e-5, atol=2e-6):
            return False
        if par is not None and not np.allclose(par, par_r, rtol=2e-5, atol=2e-6):
            return False
    return True


def test_planner_log_with_trees_matches_the_reference_planners_own_log(ob, monkeypatch, tmp_path):
    """RBTConnect / RRTConnect::outputPlannerData (RBTConnect.cpp:203-253) of the compiled reference wrote tests/golden/plan_log_*.log for
    two golden runs; the same runs here (RVS_PLAN_RNG=libc), written by planning.write_planner_log with their trees, hold the same
    `Tree: q_start` / `Tree: q_goal` / `Path:` blocks state by state (6 significant digits, the precision of the text)."""
    mod, _ = _plan_golden()
    monkeypatch.setenv("RVS_PLAN_RNG", "libc")
    for name, kind, seed in mod.LOG_RUNS:
        ref_text = open(mod.log_fixture(name, kind, seed)).read()
        o, scene, qs, qg = _oracle(ob, name)
        rb = scene[0]
        res = ob.plan_solve_batch_trees(o, mod.config(kind), qs, qg, [seed], max_path=mod.MAX_PATH, trees=8192)[0]
        assert res["success"] and res["tree_sizes"][0] + res["tree_sizes"][1] == res["states"]
        f = tmp_path / f"{name}.log"
        planning.write_planner_log(str(f), res, kind, rb.num_dofs)
        text = f.read_text()
        ours, ours_path = mod.parse_log(text)
        theirs, theirs_path = mod.parse_log(ref_text)
        assert list(ours) == list(theirs) == ["q_start", "q_goal"]
        for t in theirs:
            assert len(theirs[t]) > 5 and _same_log_states(ours[t], theirs[t]), (name, t)
        assert _same_log_states(ours_path, theirs_path)
        # everything but the space name and the wall clock is the same text up to number formatting
        head = lambda s: [l for l in s.splitlines()[:8] if not l.startswith(("Space Type", "\t Planning time"))]
        assert head(text) == head(ref_text)


@pytest.mark.gpu
def test_gpu_connect_chains_give_the_same_runs(monkeypatch):
    """The same planner batch with and without rvs_store_connect chains (RVS_PLAN_CONNECT_CHUNK): identical runs."""
    fn, qs, qg = SCENARIOS["xarm6"]
    rb, mn, mx, g, tol = fn()
    eng = engine.Engine(rb)
    eng.set_obstacles(mn, mx, g, tol)
    out = {}
    for chunk in (0, 3):
        monkeypatch.setenv("RVS_PLAN_CONNECT_CHUNK", str(chunk))
        res, st = planning.solve_batch(eng, planning.default_config(planning.RGBT_CONNECT), np.asarray(qs, np.float32), np.asarray(qg, np.float32), seeds=np.arange(24))
        out[chunk] = res
    for a, b in zip(out[0], out[3]):
        assert a["success"] and a["iterations"] == b["iterations"] and a["states"] == b["states"] and np.array_equal(a["path"], b["path"])
    eng.close()


@pytest.mark.gpu
def test_gpu_planner_returns_the_trees_of_a_run():
    fn, qs, qg = SCENARIOS["xarm6"]
    rb, mn, mx, g, tol = fn()
    eng = engine.Engine(rb)
    eng.set_obstacles(mn, mx, g, tol)
    res, _ = planning.solve_batch(eng, planning.default_config(planning.RGBT_CONNECT), np.asarray(qs, np.float32), np.asarray(qg, np.float32), seeds=[3, 4], trees=4096)
    for r in res:
        assert r["success"] and sum(r["tree_sizes"]) == r["states"]
        for t, root in zip(r["trees"], (qs, qg)):
            coord, parent = t
            assert parent[0] == -1 and (parent[1:] >= 0).all() and (parent[1:] < np.arange(1, len(parent))).all()      # parents precede children
            assert np.allclose(coord[0], root, atol=1e-6)
        # the path runs through states of the two trees
        both = np.vstack([r["trees"][0][0], r["trees"][1][0]])
        assert all((np.abs(both - p).max(axis=1) < 1e-6).any() for p in r["path"])
    eng.close()


@pytest.mark.gpu
def test_gpu_backed_planners_reproduce_the_reference_planner_runs(monkeypatch):
    """The planners on the CUDA backend against the compiled reference's runs.  A device FK that differs from the host's in the last
    ulp can flip one verdict in a run of thousands of queries; such a run diverges into another valid run, so a few are tolerated."""
    mod, gold = _plan_golden()
    monkeypatch.setenv("RVS_PLAN_RNG", "libc")
    same = total = 0
    for name in mod.SCENARIOS:
        fn, qs, qg = SCENARIOS[name]
        rb, mn, mx, g, tol = fn()
        eng = engine.Engine(rb)
        eng.set_obstacles(mn, mx, g, tol)
        qs, qg = np.asarray(qs, np.float32), np.asarray(qg, np.float32)
        for kind in mod.KINDS:
            for seed in mod.SEEDS:
                res, _ = planning.solve_batch(eng, mod.config(kind), qs, qg, seeds=[seed], max_path=mod.MAX_PATH)
                assert res[0]["success"]
                same += _same_run(res[0], gold[f"{name}_k{kind}_s{seed}_counts"], gold[f"{name}_k{kind}_s{seed}_path"], exact=False)
                total += 1
        eng.close()
    assert same >= total - 4, (same, total)
