#!/usr/bin/env python
"""profiles/rrtx_probe.py — RRTx tree growth (rvsplan_rrtx_grow) on the CUDA engine next to the same code on the CPU oracle: the shipped
configuration_rrtx.yaml (EPS_STEP 1.0, R_REWIRE 3.14, MAX_NEIGHBORS 1000) on xarm6 scenario1 and on a 200-obstacle scene, per execution mode."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes
import bench_planning
from rpmplv2_b200 import engine, planning, robots, synthetic
from oracle import oracle_binding as ob            # CPU baseline leg only

Q_GOAL = np.array([3.14159, 0.7854, -2.3562, 3.14159, 0, 3.14159], np.float32)
for name, (rb, mn, mx, g, tol), S in (("xarm6 scenario1 (5 boxes)", scenes.xarm6_scenario1(), 3000), ("xarm6 200 boxes", (robots.xarm6(),) + tuple(synthetic.sweep_obstacles(200, "D2", r_min=0.55)), 3000)):
    smp = synthetic.configs(rb, S, seed=0x7272)
    eng = engine.Engine(rb)
    sc = dict(obstacles=(mn, mx, g, tol), q_start=Q_GOAL, q_goal=Q_GOAL)
    bench_planning._prune(eng, sc)                      # the reference re-draws random obstacles that hit the start / goal (apps/CommonFunctions.h:102-110)
    mn, mx, g, tol = sc["obstacles"]
    eng.set_obstacles(mn, mx, g, tol)
    row = dict(scene=name, samples=S, obstacles=int(len(g)))
    base = None
    for mode, kw in (("one call per iteration, host neighbours", dict(one_call_per_iteration=1, device_index=0)),
                     ("three calls per iteration, host neighbours", dict(one_call_per_iteration=0, device_index=0)),
                     ("one call per iteration, device-resident tree", dict(one_call_per_iteration=1, device_index=1))):
        cfg = planning.rrtx_config(**kw)
        planning.rrtx_grow(eng, cfg, Q_GOAL, smp[:50])
        t = time.perf_counter(); out = planning.rrtx_grow(eng, cfg, Q_GOAL, smp); dt = time.perf_counter() - t
        st = out["stats"]
        row[mode] = dict(ms=1e3 * dt, us_per_iteration=1e6 * dt / S, states=st["states"], backend_calls=st["backend_calls"], edges_checked=st["edges_checked"],
                         edges_reference=st["edges_reference"], rewired=st["rewired"])
        if base is None: base = out
        else: row[mode]["same_tree"] = bool(np.array_equal(out["parent"], base["parent"]) and np.array_equal(out["cost"], base["cost"]))
    o = ob.Oracle(rb, robots.default_params()); o.set_obstacles(mn, mx, g, tol)
    t = time.perf_counter(); out_c = ob.plan_rrtx_grow(o, planning.rrtx_config(one_call_per_iteration=0), Q_GOAL, smp); dt = time.perf_counter() - t
    row["cpu oracle, 1 thread, three calls per iteration"] = dict(ms=1e3 * dt, us_per_iteration=1e6 * dt / S, states=out_c["stats"]["states"],
                                                                  same_parents=bool(len(out_c["parent"]) == len(base["parent"]) and np.array_equal(out_c["parent"], base["parent"])))
    print(json.dumps(row), flush=True)
    eng.close()
