"""The drop-in itself: the REFERENCE's own code running over base::RealVectorSpaceCUDA.

oracle/_ref/libdropin_ref.so (oracle/Makefile, built where /root/reference exists; git-ignored, travels to the GPU box prebuilt) is
the harness of oracle/rvs_ref_capi.cpp compiled with -DRVS_DROPIN: the reference's unmodified State / StateSpace / Environment / Tree /
RRTConnect / RBTConnect / RGBTConnect sources and its REAL headers, with rpmplv2_b200/cpp/RealVectorSpaceCUDA.cpp as the state space
(only the four hot virtuals isValid(q), isValid(q1,q2), computeDistance, computeDistanceUnderestimation reach the GPU, through
librvs_cuda.so's C ABI).  The golden vectors it is checked against come from the CPU build of the same harness, i.e. from the
reference's own RealVectorSpace.cpp (tests/golden/{rvs,spine,plan}_golden.npz).
"""
import ctypes
import importlib.util
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

SO = os.path.join(ROOT, "oracle", "_ref", "libdropin_ref.so")
needs_dropin = pytest.mark.skipif(not os.path.exists(SO), reason="oracle/_ref/libdropin_ref.so is built where /root/reference exists")


def _golden_module(name):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tests", "golden", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@needs_dropin
def test_dropin_library_links_the_engine_and_exports_the_harness():
    lib = ctypes.CDLL(SO)                                              # resolves librvs_cuda.so through its rpath; no GPU needed to load
    for sym in ("rvsref_create", "rvsref_set_obstacles", "rvsref_is_valid", "rvsref_is_valid_edges", "rvsref_distance", "rvsref_underestimation",
                "rvsref_extend_spines", "rvsref_plan"):
        assert hasattr(lib, sym), sym


@pytest.mark.gpu
@needs_dropin
def test_reference_loops_over_the_cuda_state_space_match_the_reference_golden():
    """isValid / computeDistance / computeDistanceUnderestimation called state by state through base::StateSpace's virtuals, State
    caches filled by RealVectorSpaceCUDA, compared with what the reference's own RealVectorSpace.cpp produced."""
    mod = _golden_module("make_rvs_golden")
    gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "rvs_golden.npz")))
    lib = ctypes.CDLL(SO)
    for name, case in mod.cases().items():
        r = mod.run_reference(lib, *case)
        G = lambda k: gold[f"{name}__{k}"]
        assert np.array_equal(r["valid"], G("valid")) and np.array_equal(r["edge"], G("edge")), name
        free = G("d") > 0
        assert np.array_equal(r["has_near"].astype(bool), free) and r["is_real"].all(), name       # nullptr exactly after a collision
        assert ((r["d"] == 0) == (G("d") == 0)).all() and np.abs(r["d"] - G("d")).max() <= 1e-4, name
        assert np.abs(r["prof"] - G("prof")).max() <= 1e-4 and np.abs(r["near"][free] - G("near")[free]).max() <= 1e-4, name
        assert (r["d"] != G("d")).mean() < 5e-3                                                     # bit-exact up to rare FK last-ulp effects
        assert int(r["root"]) == int(G("root"))
        assert np.array_equal(r["un_status"], G("un_status")) and not r["un_is_real"][G("un_status") == 0].any(), name
        acc = G("un_status") == 0
        assert np.abs(r["un_d"] - G("un_d")).max() <= 1e-4 and np.abs(r["un_prof"][acc] - G("un_prof")[acc]).max() <= 1e-4, name


@pytest.mark.gpu
@needs_dropin
def test_reference_spines_over_the_cuda_state_space_match_the_reference_golden():
    """The reference's extendSpine / extendGenSpine / extendGenSpine2 with computeDistance and computeDistanceUnderestimation served by the GPU."""
    mod = _golden_module("make_spine_golden")
    gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "spine_golden.npz")))
    lib = ctypes.CDLL(SO)
    for name, (rb, mn, mx, g, tol, thr, N) in mod.rvs.cases().items():
        h = mod.open_reference(lib, rb, mn, mx, g, tol, thr)
        free = mod.reference_free(lib, h, rb, N)
        assert np.array_equal(free, gold[f"{name}__free"])
        q_r, q_e = mod.spine_inputs(rb, N, free)
        for pi, prm in enumerate(mod.PARAMS):
            for mode in mod.MODES:
                ql, nl, dl, st = mod.run_reference_spines(lib, h, rb.num_dofs, q_r, q_e, prm, mode)
                W = lambda k: gold[f"{name}__p{pi}_m{mode}_{k}"]
                agree = (st == W("st")) & (nl == W("n"))
                assert agree.mean() > 0.99, (name, prm, mode)               # a step within an ulp of 1 / a distance within an ulp of D_CRIT may flip
                for s in np.flatnonzero(agree & (nl > 0)):                   # (extendGenSpine2 is not called on a root without clearance)
                    assert np.abs(ql[s, :nl[s]] - W("q")[s, :nl[s]]).max() <= 1e-5, (name, prm, mode, s)
                if mode != 0:
                    assert np.abs(dl[agree] - W("d")[agree]).max() <= 1e-4
        lib.rvsref_destroy(h)


@pytest.mark.gpu
@needs_dropin
def test_reference_planners_over_the_cuda_state_space_reproduce_their_cpu_runs():
    """RRTConnect / RBTConnect / RGBTConnect::solve() — the reference's code, unmodified — planning over RealVectorSpaceCUDA: every run
    succeeds, its path is collision-free, and (up to rare last-ulp verdict flips that send a run down another valid branch) it is the
    very run the reference makes over its own CPU state space."""
    mod = _golden_module("make_plan_golden")
    gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "plan_golden.npz")))
    lib = ctypes.CDLL(SO)
    same = total = 0
    for name in mod.SCENARIOS:
        h, rb, qs, qg = mod.open_reference(lib, name)
        assert h.value, "rvsref_create failed (no engine?)"
        for kind in mod.KINDS:
            for seed in mod.SEEDS:
                r = mod.run_reference(lib, h, rb.num_dofs, mod.config(kind), qs, qg, seed)
                assert not r["invalid_ends"] and r["success"], (name, kind, seed)
                assert (r["path"][0] == qs).all() and (r["path"][-1] == qg).all()
                counts, path = gold[f"{name}_k{kind}_s{seed}_counts"], gold[f"{name}_k{kind}_s{seed}_path"]
                if r["iterations"] == counts[1] and r["states"] == counts[2] and len(r["path"]) == len(path):
                    same += bool(np.abs(r["path"] - path).max() <= 1e-4)
                total += 1
        lib.rvsref_destroy(h)
    assert same >= total - 4, (same, total)
