"""RRTx's edge-batch consumers (rpmplv2_b200/cpp/planning/Rrtx.cpp, include/rvs_planning.h: rvsplan_rrtx_grow)
against trees grown by the reference's OWN RRTx members (findNeighbors, chooseParent, rewireNeighbors, propagateCostChanges, updatePath,
computePath, shrinkingBallRadius; src/planners/rrtx/RRTx.cpp compiled unmodified, tests/golden/make_rrtx_golden.py ->
tests/golden/rrtx_golden.npz).  Bar: identical trees — coordinates, parents and costs bit for bit — although every candidate edge of an
iteration is answered by ONE batched query instead of one call per edge."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import ROOT
from rpmplv2_b200 import engine, planning, robots

spec = importlib.util.spec_from_file_location("make_rrtx_golden", os.path.join(ROOT, "tests", "golden", "make_rrtx_golden.py"))
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)
GOLD = dict(np.load(os.path.join(ROOT, "tests", "golden", "rrtx_golden.npz")))
CASES = [(name, v) for name in mod.SCENARIOS for v in mod.VARIANTS]


def same_tree(out, name, tag, exact=True):
    g = {k: GOLD[f"{name}_{tag}_{k}"] for k in ("coord", "parent", "cost", "new_index", "path")}
    if not (len(out["parent"]) == len(g["parent"]) and np.array_equal(out["new_index"], g["new_index"]) and np.array_equal(out["parent"], g["parent"])):
        return False
    if exact:
        return np.array_equal(out["coord"], g["coord"]) and np.array_equal(out["cost"], g["cost"])
    return bool(np.abs(out["coord"] - g["coord"]).max() <= 1e-4 and np.abs(out["cost"] - g["cost"]).max() <= 1e-3)


def oracle_for(ob, name):
    fn, qs, qg = mod.SCENARIOS[name]
    rb, mn, mx, g, tol = fn()
    o = ob.Oracle(rb, robots.default_params())
    o.set_obstacles(mn, mx, g, tol)
    return o, rb, np.asarray(qg, np.float32)


@pytest.mark.parametrize("name,variant", CASES, ids=lambda c: c if isinstance(c, str) else c[0])
def test_oracle_backed_rrtx_reproduces_the_reference_tree(ob, name, variant):
    tag, eps, r, cap, S = variant
    o, rb, qg = oracle_for(ob, name)
    smp = mod.samples_for(rb, name, tag, S)
    start, invalid = GOLD[f"{name}_{tag}_path_start"]
    for one_call in (1, 0):
        out = ob.plan_rrtx_grow(o, mod.config(eps, r, cap, one_call_per_iteration=one_call), qg, smp, path_start=int(start))
        assert same_tree(out, name, tag), (name, tag, one_call)
        assert np.array_equal(out["path"], GOLD[f"{name}_{tag}_path"]) and out["path_invalid"] == bool(invalid)
        st = out["stats"]
        assert st["states"] == len(out["parent"]) and st["iterations"] == S and st["trapped"] == int((out["new_index"] < 0).sum())
        # batching asks a superset of the reference's one-at-a-time checks, in far fewer calls
        assert st["edges_checked"] >= st["edges_reference"]
        if one_call and not o.robot.self_collision_checking:
            assert st["backend_calls"] == S + 1                                   # one per iteration + the path walk
        if not one_call and not o.robot.self_collision_checking:
            assert st["backend_calls"] <= 3 * S + 1


def test_rrtx_radius_is_the_reference_shrinking_ball(ob):
    """The radii the compiled RRTx::shrinkingBallRadius returned while the golden 'ball' trees grew are what RrtxTree uses: the trees
    only match if every radius does; here the values themselves are looked at once (gamma_rrt, eta of RRTx's constructor)."""
    for name in mod.SCENARIOS:
        radii = GOLD[f"{name}_ball_radii"]
        n = GOLD[f"{name}_ball_coord"].shape[1]
        assert radii[0] == 0.0                                                    # log(1) / 1
        assert (np.diff(radii[5:]) <= 0).all() and radii.max() <= np.float32(np.sqrt(n))


def test_oracle_backed_rrtx_against_the_compiled_reference_on_fresh_samples(ob):
    so = os.path.join(ROOT, "oracle", "_ref", "librvs_ref.so")
    if not os.path.exists(so) or not os.path.isdir("/root/reference/data"):
        pytest.skip("oracle/_ref is built where /root/reference exists")
    import refharness
    lib = refharness.load()
    if not hasattr(lib, "rvsref_rrtx_grow"):
        pytest.skip("oracle/_ref/librvs_ref.so predates the RRTx harness")
    from rpmplv2_b200 import synthetic
    for name in mod.SCENARIOS:
        h, rb, qs, qg = mod.open_reference(lib, name)
        o, _, _ = oracle_for(ob, name)
        for vi, (eps, r, cap) in enumerate([(0.7, 1.5, 40), (0.2, 0.0, 1000)]):
            smp = synthetic.configs(rb, 500, seed=4242 + vi)
            want = mod.run_reference(lib, h, rb.num_dofs, eps, r, cap, qs, qg, smp)
            out = ob.plan_rrtx_grow(o, mod.config(eps, r, cap), qg, smp)
            assert np.array_equal(out["parent"], want["parent"]) and np.array_equal(out["cost"], want["cost"]) and np.array_equal(out["coord"], want["coord"])
        lib.rvsref_destroy(h)


@pytest.mark.gpu
@pytest.mark.parametrize("name,variant", CASES, ids=lambda c: c if isinstance(c, str) else c[0])
def test_gpu_backed_rrtx_reproduces_the_reference_tree(name, variant):
    """rvsplan_rrtx_grow on the CUDA engine.  The device's FK may differ from the host's in the last ulp, which can flip one edge verdict
    in tens of thousands and send the tree down another (valid) history: parents must match, coordinates / costs to 1e-4 / 1e-3."""
    tag, eps, r, cap, S = variant
    fn, qs, qg = mod.SCENARIOS[name]
    rb, mn, mx, g, tol = fn()
    eng = engine.Engine(rb)
    eng.set_obstacles(mn, mx, g, tol)
    smp = mod.samples_for(rb, name, tag, S)
    start, invalid = GOLD[f"{name}_{tag}_path_start"]
    outs = []
    for one_call, dev in ((1, 0), (0, 0), (1, 1)):
        launches = eng.launch_count()
        out = planning.rrtx_grow(eng, mod.config(eps, r, cap, one_call_per_iteration=one_call, device_index=dev), np.asarray(qg, np.float32), smp,
                                 path_start=int(start))
        assert eng.launch_count() > launches
        assert same_tree(out, name, tag, exact=False), (name, tag, one_call, dev)
        assert np.array_equal(out["path"], GOLD[f"{name}_{tag}_path"]) and out["path_invalid"] == bool(invalid)
        outs.append(out)
    # the three execution modes are the same tree bit for bit (same engine, same verdicts; device and host neighbour queries agree)
    for o2 in outs[1:]:
        assert np.array_equal(o2["parent"], outs[0]["parent"]) and np.array_equal(o2["cost"], outs[0]["cost"]) and np.array_equal(o2["coord"], outs[0]["coord"])
    eng.close()
