#!/usr/bin/env python
"""profiles/tree_probe.py — pattern trees (rpmplv2_b200/pattern_tree.py): time inside the engine calls (two per layer) on the B200 next
to the same calls answered by the CPU oracle port on one host thread (what the reference does node by node).  Evidence only."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes  # noqa: E402
from oracle import oracle_binding as ob  # noqa: E402
from oracle.oracle_space import OracleSpace  # noqa: E402
from rpmplv2_b200 import engine, pattern_tree, robots  # noqa: E402


class Timed:
    def __init__(self, space):
        self.s, self.t, self.calls = space, 0.0, 0

    def store_reserve(self, c):
        self.s.store_reserve(c)

    def _timed(self, fn, *a, **k):
        t0 = time.perf_counter()
        r = fn(*a, **k)
        self.t += time.perf_counter() - t0
        self.calls += 1
        return r

    def store_distance(self, *a, **k):
        return self._timed(self.s.store_distance, *a, **k)

    def store_extend_spines(self, *a, **k):
        return self._timed(self.s.store_extend_spines, *a, **k)


if __name__ == "__main__":
    cases = [("xarm6 scenario1", scenes.xarm6_scenario1(), [0, 0.7854, -2.3562, 0, 0, 0], 3),
             ("xarm6 + 1000 D2 boxes", scenes.xarm6_scene(1000, "D2", 0.9), [0, 0.7854, -2.3562, 0, 0, 0], 3),
             ("spatial_10dof scenario1", scenes.spatial10_scenario1(), [0, 1.57, 0, 0, 0, 0, 0, 0, 0, 0], 3)]
    for name, (rb, mn, mx, g, tol), q, layers in cases:
        q = np.asarray(q, np.float32)
        eng = engine.Engine(rb)
        eng.set_obstacles(mn, mx, g, tol)
        o = ob.Oracle(rb, robots.default_params())
        o.set_obstacles(mn, mx, g, tol)
        row = {"scene": name, "tree_layers": layers, "obstacles": int(len(g))}
        for arm, space in (("gpu", eng), ("cpu_port_1_thread", OracleSpace(o))):
            best = None
            for rep in range(3 if arm == "gpu" else 1):
                t = Timed(space)
                t0 = time.perf_counter()
                tree = pattern_tree.generate_local_tree(t, rb, q, layers)
                total = time.perf_counter() - t0
                if best is None or t.t < best[0]:
                    best = (t.t, t.calls, total, len(tree["coord"]))
            row[arm] = {"ms_in_space_calls": round(1e3 * best[0], 3), "space_calls": best[1], "ms_total_with_python_host": round(1e3 * best[2], 1), "nodes": best[3]}
        row["reference_calls_for_the_same_tree"] = "1 computeDistance + up to 2n extendGenSpine per expanded node"
        print(json.dumps(row), flush=True)
        eng.close()
