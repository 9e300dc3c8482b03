#!/usr/bin/env python
"""tests/golden/make_plan_golden.py — golden planner runs produced by the REFERENCE's own
planning::rrt::RRTConnect / planning::rbt::RBTConnect / planning::rbt::RGBTConnect::solve() (src/planners/rrt/RRTConnect.cpp:41-85,
src/planners/rbt/RBTConnect.cpp:14-64, RGBTConnect.cpp:14-69) with base::Tree and base::RealVectorSpace underneath, all compiled
unmodified against oracle/eigen_shim + oracle/env_shim (oracle/Makefile target `ref` -> oracle/_ref/librvs_ref.so, entry rvsref_plan).
The planners' only randomness is Eigen::VectorXf::Random — in the work-alike one std::rand() per coefficient — so every run is a
function of srand(seed); a planner that draws from the same stream (RVS_PLAN_RNG=libc) must reproduce it state by state.
The harness robot reports no self-collision and answers computeSkeleton / computeEnclosingRadii with the oracle's restatements.
Run in the build container (needs /root/reference); writes tests/golden/plan_golden.npz."""
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
from rpmplv2_b200 import planning  # noqa: E402

f32p, u8p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint8)
P = lambda a, t: a.ctypes.data_as(t)
MAX_PATH = 4096
# q_start / q_goal as in tests/test_planners.py (the reference's shipped scenario files)
SCENARIOS = {
    "planar2": (scenes.planar2_scenario_test, [-2, -2.5], [2, 2.5]),
    "xarm6": (scenes.xarm6_scenario1, [0, 0.7854, -2.3562, 0, 0, 0], [3.14159, 0.7854, -2.3562, 3.14159, 0, 3.14159]),
    "spatial10": (scenes.spatial10_scenario1, [0, 1.57, 0, 0, 0, 0, 0, 0, 0, 0], [3.14, 1.57, 0, 0, 0, 0, 0, 0, 0, 0]),
}
KINDS = (planning.RRT_CONNECT, planning.RBT_CONNECT, planning.RGBT_CONNECT)
SEEDS = (1, 2, 3)


def config(kind, **over):
    return planning.default_config(kind, max_planning_time=600.0, max_num_iter=20000, **over)


def open_reference(lib, name):
    fn, qs, qg = SCENARIOS[name]
    rb, mn, mx, g, tol = fn()
    h, _ = refharness.create(lib, rb, 1e-4)      # the reference's own robot class; xArm6 plans with self_collision_checking = true
    lib.rvsref_set_self_collision(h, int(rb.self_collision_checking))
    mn, mx, g, tol = [np.ascontiguousarray(a) for a in (mn, mx, g, tol)]
    lib.rvsref_set_obstacles(h, P(mn, f32p), P(mx, f32p), P(g, u8p), P(tol, f32p), len(g))
    return h, rb, np.asarray(qs, np.float32), np.asarray(qg, np.float32)


def run_reference(lib, h, n, cfg, qs, qg, seed):
    suc, pl, it, st, ts = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int64(), ctypes.c_int64(), (ctypes.c_int64 * 2)()
    path = np.zeros((MAX_PATH, n), np.float32)
    rc = lib.rvsref_plan(h, cfg.kind, P(qs, f32p), P(qg, f32p), ctypes.c_uint32(seed), ctypes.c_int64(cfg.max_num_iter), ctypes.c_int64(cfg.max_num_states),
                         ctypes.c_float(cfg.max_planning_time), ctypes.c_float(cfg.eps_step), cfg.max_extension_steps, ctypes.c_float(cfg.delta),
                         cfg.num_spines, cfg.spine.num_layers, cfg.spine.num_iter_spine, cfg.spine.use_expanded_bubble, ctypes.c_float(cfg.spine.d_crit),
                         ctypes.byref(suc), ctypes.byref(it), ctypes.byref(st), ctypes.byref(pl), P(path, f32p), MAX_PATH, ts)
    if rc == 2:
        return dict(invalid_ends=True)
    assert rc == 0 and pl.value <= MAX_PATH
    return dict(invalid_ends=False, success=bool(suc.value), iterations=it.value, states=st.value, path=path[:pl.value].copy(), tree_sizes=(ts[0], ts[1]))


LOG_RUNS = (("xarm6", planning.RGBT_CONNECT, 1), ("planar2", planning.RRT_CONNECT, 2))     # planner logs kept as text fixtures


def log_fixture(name, kind, seed):
    return os.path.join(ROOT, "tests", "golden", f"plan_log_{name}_k{kind}_s{seed}.log")


def parse_log(text):
    """-> ({tree name: [(q, parent q or None), ...]}, path [(q, parent q or None), ...]) from a planner log (RBTConnect::outputPlannerData's format)."""
    def state(line):
        a, b = line.split("; parent q: ")
        q = [float(x) for x in a[len("q: ("):-1].split()]
        return q, None if b.strip() == "NONE" else [float(x) for x in b.strip()[1:-1].split()]
    trees, path, cur = {}, [], None
    for line in text.splitlines():
        if line.startswith("Tree: "):
            cur = trees.setdefault(line[len("Tree: "):].strip(), [])
        elif line.startswith("Path:"):
            cur = path
        elif line.startswith("q: (") and cur is not None:
            cur.append(state(line))
        elif line.startswith("-----"):
            break
    return trees, path


if __name__ == "__main__":
    lib = refharness.load()
    data = {}
    for name in SCENARIOS:
        h, rb, qs, qg = open_reference(lib, name)
        for kind in KINDS:
            for seed in SEEDS:
                r = run_reference(lib, h, rb.num_dofs, config(kind), qs, qg, seed)
                print(name, kind, seed, r["success"], r["iterations"], r["states"], len(r["path"]))
                data[f"{name}_k{kind}_s{seed}_counts"] = np.array([r["success"], r["iterations"], r["states"], r["tree_sizes"][0], r["tree_sizes"][1]], np.int64)
                data[f"{name}_k{kind}_s{seed}_path"] = r["path"]
        for lname, lkind, lseed in LOG_RUNS:
            if lname == name:                                            # the reference's own log of that run (outputPlannerData)
                lib.rvsref_set_plan_log(log_fixture(lname, lkind, lseed).encode())
                run_reference(lib, h, rb.num_dofs, config(lkind), qs, qg, lseed)
                lib.rvsref_set_plan_log(None)
                import re
                text = open(log_fixture(lname, lkind, lseed)).read()                # the wall-clock line is the only run-dependent one
                open(log_fixture(lname, lkind, lseed), "w").write(re.sub(r"Planning time \[s\]:    .*", "Planning time [s]:    0", text))
        lib.rvsref_destroy(h)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "plan_golden.npz"), **data)
