#!/usr/bin/env python
"""tests/golden/make_rrtx_golden.py — golden RRTx trees produced by the REFERENCE's own planning::rrtx::RRTx
members (src/planners/rrtx/RRTx.cpp: findNeighbors :424-452, chooseParent :454-491, rewireNeighbors :493-529, propagateCostChanges
:531-548, updatePath :550-572, computePath :575-592, shrinkingBallRadius :636-639 and the constants of its constructor :71-75;
RRTConnect::extend, base::Tree::getNearestState / upgradeTree underneath), compiled unmodified into oracle/_ref/librvs_ref.so and driven
in the order of RRTx::solve's tree-growth block (:157-190) by rvsref_rrtx_grow on given random states.
Run in the build container (needs /root/reference); writes tests/golden/rrtx_golden.npz."""
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
from rpmplv2_b200 import planning, synthetic  # noqa: E402

f32p, u8p, i32p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_int32)
P = lambda a, t: a.ctypes.data_as(t)
# q_start (only has to be valid: RRTx's constructor checks it) / q_goal (the root of RRTx's tree) as in tests/test_planners.py
SCENARIOS = {
    "planar2": (scenes.planar2_scenario_test, [-2, -2.5], [2, 2.5]),
    "xarm6": (scenes.xarm6_scenario1, [0, 0.7854, -2.3562, 0, 0, 0], [3.14159, 0.7854, -2.3562, 3.14159, 0, 3.14159]),
    "spatial10": (scenes.spatial10_scenario1, [0, 1.57, 0, 0, 0, 0, 0, 0, 0, 0], [3.14, 1.57, 0, 0, 0, 0, 0, 0, 0, 0]),
}
# (tag, EPS_STEP, R_REWIRE, MAX_NEIGHBORS, samples): the shipped configuration_rrtx.yaml, a shrinking ball, and a neighbour cap that bites
VARIANTS = (("yaml", 1.0, 3.14, 1000, 700), ("ball", 0.3, 0.0, 1000, 900), ("cap", 0.5, 2.0, 12, 700))


def samples_for(rb, name, tag, S):
    seed = 0x52525458 + 977 * sorted(SCENARIOS).index(name) + 31 * [v[0] for v in VARIANTS].index(tag)
    return synthetic.configs(rb, S, seed=seed)


def config(eps_step, r_rewire, max_neighbors, **over):
    return planning.rrtx_config(eps_step=eps_step, r_rewire=r_rewire, max_neighbors=max_neighbors, **over)


def open_reference(lib, name):
    fn, qs, qg = SCENARIOS[name]
    rb, mn, mx, g, tol = fn()
    h, _ = refharness.create(lib, rb, 1e-4)      # the reference's own robot class; xArm6 runs with self_collision_checking = true
    lib.rvsref_set_self_collision(h, int(rb.self_collision_checking))
    refharness.set_obstacles(lib, h, mn, mx, g, tol)
    return h, rb, np.asarray(qs, np.float32), np.asarray(qg, np.float32)


def run_reference(lib, h, n, eps_step, r_rewire, max_neighbors, qs, qg, samples, path_start=-1):
    lib.rvsref_rrtx_grow.restype = ctypes.c_int64
    samples = np.ascontiguousarray(samples, np.float32)
    S = len(samples)
    cap = S + 1
    coord, parent, cost = np.zeros((cap, n), np.float32), np.zeros(cap, np.int32), np.zeros(cap, np.float32)
    new_index, radii, path = np.zeros(S, np.int32), np.zeros(S, np.float32), np.zeros(cap, np.int32)
    path_len, path_invalid = ctypes.c_int32(0), ctypes.c_int32(0)
    count = lib.rvsref_rrtx_grow(h, ctypes.c_float(eps_step), ctypes.c_float(r_rewire), int(max_neighbors), P(qs, f32p), P(qg, f32p), P(samples, f32p), S,
                                 ctypes.c_int64(cap), P(coord, f32p), P(parent, i32p), P(cost, f32p), P(new_index, i32p), P(radii, f32p), int(path_start),
                                 P(path, i32p), ctypes.byref(path_len), ctypes.byref(path_invalid))
    assert count > 0, count
    return dict(coord=coord[:count].copy(), parent=parent[:count].copy(), cost=cost[:count].copy(), new_index=new_index, radii=radii,
                path=path[:path_len.value].copy(), path_invalid=bool(path_invalid.value))


if __name__ == "__main__":
    lib = refharness.load()
    data = {}
    for name in SCENARIOS:
        h, rb, qs, qg = open_reference(lib, name)
        for tag, eps, r, cap, S in VARIANTS:
            smp = samples_for(rb, name, tag, S)
            r0 = run_reference(lib, h, rb.num_dofs, eps, r, cap, qs, qg, smp)
            start = int(np.argmax(r0["cost"] * (r0["parent"] >= 0)))           # the costliest connected state: the longest chain to the root
            res = run_reference(lib, h, rb.num_dofs, eps, r, cap, qs, qg, smp, path_start=start)
            assert np.array_equal(res["parent"], r0["parent"]) and np.array_equal(res["cost"], r0["cost"])
            print(name, tag, "states", len(res["parent"]), "orphans", int((res["parent"][1:] < 0).sum()), "path", len(res["path"]), res["path_invalid"],
                  "reparented", int((res["parent"][1:] != np.arange(len(res["parent"]) - 1)).sum()))
            for k in ("coord", "parent", "cost", "new_index", "radii", "path"):
                data[f"{name}_{tag}_{k}"] = res[k]
            data[f"{name}_{tag}_path_start"] = np.array([start, int(res["path_invalid"])], np.int32)
        lib.rvsref_destroy(h)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "rrtx_golden.npz"), **data)
