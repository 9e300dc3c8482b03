#!/usr/bin/env python
"""tests/golden/make_tree_golden.py — golden pattern trees (SURVEY.md §8 row f-1 "PatternTree layers"; the trees behind the rows of
apps/generate_dataset.cpp) produced by the REFERENCE's own planning::rbt::PatternTree::generateLocalTree
(src/planners/rbt/PatternTree.cpp:20-91), compiled unmodified into oracle/_ref/librvs_ref.so (entry rvsref_pattern_tree).
Run in the build container (needs /root/reference); writes tests/golden/tree_golden.npz."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refharness  # noqa: E402
import make_plan_golden as mp  # noqa: E402

f32p, u8p, i32p = mp.f32p, mp.u8p, ctypes.POINTER(ctypes.c_int32)
P = mp.P
# (scenario, tree layers, (NUM_LAYERS, NUM_ITER_SPINE, USE_EXPANDED_BUBBLE, D_CRIT)); the root is the scenario's q_start, then q_goal
CASES = [("planar2", 3, (5, 5, 1, 0.03)), ("planar2", 2, (3, 1, 0, 0.05)), ("xarm6", 2, (5, 5, 1, 0.03)), ("spatial10", 2, (5, 1, 1, 0.03))]
DELTA = 3.14159


def run_reference(lib, h, n, q_root, tree_layers, prm, max_nodes=20000):
    lib.rvsref_pattern_tree.restype = ctypes.c_int64
    coord, parent = np.zeros((max_nodes, n), np.float32), np.zeros(max_nodes, np.int32)
    d_c, prof, real = np.zeros(max_nodes, np.float32), np.zeros((max_nodes, n), np.float32), np.zeros(max_nodes, np.uint8)
    count = lib.rvsref_pattern_tree(h, P(np.ascontiguousarray(q_root, np.float32), f32p), tree_layers, prm[0], prm[1], prm[2], ctypes.c_float(prm[3]),
                                    ctypes.c_float(DELTA), ctypes.c_int64(max_nodes), P(coord, f32p), P(parent, i32p), P(d_c, f32p), P(prof, f32p), P(real, u8p))
    assert count > 0
    return dict(coord=coord[:count].copy(), parent=parent[:count].copy(), d_c=d_c[:count].copy(), profile=prof[:count].copy(), is_real=real[:count].copy())


if __name__ == "__main__":
    lib = refharness.load()
    data = {}
    for k, (name, layers, prm) in enumerate(CASES):
        h, rb, qs, qg = mp.open_reference(lib, name)
        for tag, q in (("start", qs), ("goal", qg)):
            r = run_reference(lib, h, rb.num_dofs, q, layers, prm)
            print(name, layers, prm, tag, "nodes", len(r["coord"]))
            for key, v in r.items():
                data[f"c{k}_{tag}_{key}"] = v
        lib.rvsref_destroy(h)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tree_golden.npz"), **data)
