#!/usr/bin/env python
"""tests/golden/make_spine_golden.py — golden vectors for the bur spines, produced by the REFERENCE's own
planning::rbt::RBTConnect::extendSpine and RGBTConnect::extendGenSpine / extendGenSpine2 (src/planners/rbt/RBTConnect.cpp:85-150,
RGBTConnect.cpp:73-118), compiled unmodified together with RealVectorSpace.cpp against oracle/eigen_shim + oracle/env_shim
(oracle/Makefile target `ref` -> oracle/_ref/librvs_ref.so).  The harness robot answers computeSkeleton / computeEnclosingRadii with
the oracle's restatements (the robot classes need KDL).  Run in the build container (needs /root/reference); writes
tests/golden/spine_golden.npz."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refharness  # noqa: E402
import make_rvs_golden as rvs  # noqa: E402
from rpmplv2_b200 import synthetic  # noqa: E402

f32p, u8p, i32p = rvs.f32p, rvs.u8p, ctypes.POINTER(ctypes.c_int32)
P = rvs.P
# (NUM_LAYERS, NUM_ITER_SPINE, USE_EXPANDED_BUBBLE, D_CRIT): DRGBT's setting, the RBT/RGBT-Connect defaults (RBTConnectConfig.cpp,
# RGBTConnectConfig.cpp), and the plain-bubble branch
PARAMS = [(5, 1, 1, 0.05), (5, 5, 1, 0.03), (3, 2, 0, 0.03)]
MODES = {0: "extendSpine", 1: "extendGenSpine", 2: "extendGenSpine2"}
SPINES_PER_ROOT, MAX_ROOTS = 4, 40


def spine_inputs(rb, N, root_free):
    """Roots: the first collision-free configurations of make_rvs_golden.inputs (every planner root has d_c > 0 except the Trapped
    case, which gets ONE colliding root at the end when the scene has one); targets: random, every other one pulled close to its root
    so that `Reached` occurs."""
    q, _, _ = rvs.inputs(rb, N)
    roots = q[root_free][:MAX_ROOTS]
    q_r = np.repeat(roots, SPINES_PER_ROOT, axis=0)
    if (~root_free).any():
        q_r = np.concatenate([q_r, q[~root_free][:1]])
    S = len(q_r)
    q_e = synthetic.configs(rb, S, seed=0x33)
    q_e[::2] = (q_r[::2] + np.float32(0.15) * (q_e[::2] - q_r[::2])).astype(np.float32)
    return np.ascontiguousarray(q_r), np.ascontiguousarray(q_e)


def open_reference(lib, rb, mn, mx, g, tol, thr):
    # the reference's own robot class where it ships one: its computeEnclosingRadii and — xArm6, self_collision_checking = true as in
    # every shipped xarm6 scenario — its checkSelfCollision(q_temp, q_new) inside extendSpine (RBTConnect.cpp:129-131)
    h, _ = refharness.create(lib, rb, thr)
    lib.rvsref_set_self_collision(h, int(rb.self_collision_checking))
    mn, mx, g, tol = [np.ascontiguousarray(a) for a in (mn, mx, g, tol)]
    lib.rvsref_set_obstacles(h, P(mn, f32p), P(mx, f32p), P(g, u8p), P(tol, f32p), len(g))
    return h


def reference_free(lib, h, rb, N):
    q, _, _ = rvs.inputs(rb, N)
    v = np.zeros(N, np.uint8)
    lib.rvsref_is_valid(h, P(q, f32p), ctypes.c_int64(N), P(v, u8p))
    return v.astype(bool)


def run_reference_spines(lib, h, n, q_r, q_e, prm, mode):
    layers, iters, expanded, d_crit = prm
    S = len(q_r)
    ql, nl = np.zeros((S, layers, n), np.float32), np.zeros(S, np.int32)
    dl, st = np.zeros(S, np.float32), np.zeros(S, np.uint8)
    assert lib.rvsref_extend_spines(h, P(q_r, f32p), P(q_e, f32p), S, layers, iters, expanded, ctypes.c_float(d_crit), mode,
                                    P(ql, f32p), P(nl, i32p), P(dl, f32p), P(st, u8p)) == 0
    return ql, nl, dl, st


if __name__ == "__main__":
    lib = refharness.load()
    data = {}
    for name, (rb, mn, mx, g, tol, thr, N) in rvs.cases().items():
        h = open_reference(lib, rb, mn, mx, g, tol, thr)
        free = reference_free(lib, h, rb, N)
        data[f"{name}__free"] = free
        q_r, q_e = spine_inputs(rb, N, free)
        for pi, prm in enumerate(PARAMS):
            for mode in MODES:
                ql, nl, dl, st = run_reference_spines(lib, h, rb.num_dofs, q_r, q_e, prm, mode)
                for k, v in dict(q=ql, n=nl, d=dl, st=st).items():
                    data[f"{name}__p{pi}_m{mode}_{k}"] = v
        lib.rvsref_destroy(h)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "spine_golden.npz"), **data)
    print("wrote", len(rvs.cases()), "cases x", len(PARAMS), "parameter sets x", len(MODES), "modes")
