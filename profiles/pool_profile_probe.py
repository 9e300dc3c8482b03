import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import bench_planning
from rpmplv2_b200 import planning, robots, engine
os.environ["RVS_PLAN_PROFILE"] = "1"
for sc in bench_planning.scenarios()[:3]:
    rb = sc["robot"]; eng = engine.Engine(rb); mn, mx, g, tol = sc["obstacles"]; eng.set_obstacles(mn, mx, g, tol)
    cfg = planning.default_config(sc["planner"], max_planning_time=30.0)
    for P in (1, 64, 1024):
        planning.solve_batch(eng, cfg, sc["q_start"], sc["q_goal"], seeds=np.arange(2))
        t = time.perf_counter(); res, st = planning.solve_batch(eng, cfg, sc["q_start"], sc["q_goal"], seeds=np.arange(P)); dt = time.perf_counter() - t
        print(sc["name"], "P", P, "wall ms %.1f per run %.3f" % (1e3 * dt, 1e3 * dt / P), st, flush=True)
    eng.close()
