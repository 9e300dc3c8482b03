#!/usr/bin/env python
"""tests/golden/make_rvs_golden.py — golden vectors for the LOOPS of the hot path (isValid(q), isValid(q1,q2), computeDistance with
profile / nearest points, computeDistanceUnderestimation), produced by the REFERENCE's own RealVectorSpace.cpp compiled here against
oracle/eigen_shim + oracle/env_shim (oracle/Makefile target `ref` -> oracle/_ref/librvs_ref.so; the harness robot's computeSkeleton is
the oracle's FK restatement).  Run in the build container (needs /root/reference); writes tests/golden/rvs_golden.npz."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refharness  # noqa: E402
import scenes  # noqa: E402
from rpmplv2_b200 import robots, synthetic  # noqa: E402

f32p, u8p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint8)
P = lambda a, t: a.ctypes.data_as(t)


def cases():
    """name -> (robot, aabb_min, aabb_max, is_ground, min_dist_tol, equality_threshold, N)"""
    out = {}
    out["xarm6_m30"] = scenes.xarm6_scene(30, "D2", 0.45) + (1e-4, 160)
    out["xarm6_m200_d1"] = scenes.xarm6_scene(200, "D1", 0.2) + (1e-4, 120)
    out["planar2_scenario_test"] = scenes.planar2_scenario_test() + (1e-4, 200)
    out["spatial10_scenario1"] = scenes.spatial10_scenario1() + (1e-4, 160)
    out["spatial22_chain"] = scenes.chain_scene(robots.spatial(22), 12) + (1e-6, 100)
    rb, mn, mx, g, tol = scenes.xarm6_scene(60, "D2", 0.4)
    tol = np.random.default_rng(5).choice(np.array([0.02, 0.1, 0.4, 1000.0], np.float32), len(tol)).astype(np.float32)
    out["xarm6_tol_mix_thr005"] = (rb, mn, mx, g, tol, 0.05, 160)
    return out


def inputs(rb, N):
    q = synthetic.configs(rb, N, seed=0xAB)
    q1, q2 = synthetic.edges(rb, N, 0.3, seed=0x17)
    return q, q1, q2


def run_reference(lib, rb, mn, mx, g, tol, thr, N):
    h, _ = refharness.create(lib, rb, thr)       # the reference's own robot class (FK through the KDL work-alike) where it ships one
    mn, mx, g, tol = [np.ascontiguousarray(a) for a in (mn, mx, g, tol)]
    M, n = len(g), rb.num_dofs
    lib.rvsref_set_obstacles(h, P(mn, f32p), P(mx, f32p), P(g, u8p), P(tol, f32p), M)
    q, q1, q2 = inputs(rb, N)
    r = dict(valid=np.zeros(N, np.uint8), edge=np.zeros(N, np.uint8), d=np.zeros(N, np.float32), prof=np.zeros((N, n), np.float32),
             near=np.zeros((N, M, n, 6), np.float32), has_near=np.zeros(N, np.uint8), is_real=np.zeros(N, np.uint8))
    lib.rvsref_is_valid(h, P(q, f32p), ctypes.c_int64(N), P(r["valid"], u8p))
    lib.rvsref_is_valid_edges(h, P(q1, f32p), P(q2, f32p), ctypes.c_int64(N), P(r["edge"], u8p))
    assert lib.rvsref_distance(h, P(q, f32p), ctypes.c_int64(N), P(r["d"], f32p), P(r["prof"], f32p), P(r["near"], f32p), P(r["has_near"], u8p), P(r["is_real"], u8p)) == 0
    root = int(np.argmax(r["d"]))
    planes = np.ascontiguousarray(r["near"][root])
    r.update(root=np.int32(root), un_d=np.zeros(N, np.float32), un_prof=np.zeros((N, n), np.float32), un_status=np.zeros(N, np.uint8), un_is_real=np.zeros(N, np.uint8))
    lib.rvsref_underestimation(h, P(q, f32p), ctypes.c_int64(N), P(planes, f32p), P(r["un_d"], f32p), P(r["un_prof"], f32p), P(r["un_status"], u8p), P(r["un_is_real"], u8p))
    lib.rvsref_destroy(h)
    return r


if __name__ == "__main__":
    lib = refharness.load()
    data = {}
    for name, (rb, mn, mx, g, tol, thr, N) in cases().items():
        r = run_reference(lib, rb, mn, mx, g, tol, thr, N)
        r["near"] = r["near"].astype(np.float32)
        for k, v in r.items():
            data[f"{name}__{k}"] = v
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "rvs_golden.npz"), **data)
    print("wrote", len(cases()), "cases")
