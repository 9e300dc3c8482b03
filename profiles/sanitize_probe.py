#!/usr/bin/env python
"""profiles/sanitize_probe.py — one small call of every kernel family, meant to run under compute-sanitizer (memcheck / racecheck)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes
from rpmplv2_b200 import drgbt, engine, planning, robots, synthetic
from rpmplv2_b200.engine import env_params

for rb, (mn, mx, g, tol) in ((robots.xarm6(), synthetic.sweep_obstacles(300, "D2", r_min=0.5)), (robots.spatial(22), scenes.chain_scene(robots.spatial(22), 40)[1:])):
    eng = engine.Engine(rb)
    eng.set_obstacles(mn, mx, g, tol)
    q = synthetic.configs(rb, 6000)
    for mode in (robots.DIST_MIN, robots.DIST_PROFILE):
        eng.distance(q, mode)                       # work-queue kernel
        eng.distance(q[:100], mode)                 # small-batch kernel
    eng.distance(q[:200], robots.DIST_FULL)
    eng.set_culling(False); eng.distance(q[:500], robots.DIST_FULL); eng.is_valid(q[:500]); eng.set_culling(True)
    eng.is_valid(q); eng.is_valid(q[:40]); eng.is_valid(q[:3000])
    q1, q2 = synthetic.edges(rb, 800, 0.1)
    eng.is_valid_edges(q1, q2); eng.is_valid_edges(q1[:20], q2[:20])
    eng.distance(synthetic.configs(rb, 8300 if rb.num_dofs > 6 else 2500, seed=9), robots.DIST_FULL)     # three-launch split (rvs_distance_full.cuh): rings, obstacle ranges, two tiles / > 48 KB of shared memory
    d, p, near = eng.distance(q[:64], robots.DIST_FULL)
    k = int(np.argmax(d))
    eng.underestimation(q[:64], near[k:k + 1], np.zeros(64, np.int32))
    eng.skeleton(q[:64]); eng.enclosing_radii(q[:64]); eng.self_collision_capsules(q[:64])
    g_ = drgbt.generate_gbur(eng, q[k], q[:50])
    eng.extend_spines(q[k], p[k], float(d[k]), near[k], q[:30])
    sc = scenes.dynamic_scene(rb, 5, 3, r_lo=0.5, r_hi=0.9)
    eng.env_set(env_params("light_directions", ws_center=scenes.WS_CENTER, ws_radius=1.0, base_radius=0.3, robot_max_vel=1.5), **sc)
    eng.env_step(0.01, 7)
    eng.motion_validity(q1[:50], q2[:50], 10, 0.005, commit_edge=3)
    eng.distance(q[:5000], robots.DIST_PROFILE)
    # round 2: random_directions (libc stream on the device), self-collision edges, device-resident trees
    eng.env_set(env_params("random_directions", ws_center=scenes.WS_CENTER, ws_radius=1.0, base_radius=0.15, robot_max_vel=3.0), **sc)
    eng.env_set_random_directions(np.full((len(sc["max_vel"]), 3), 0.2, np.float32), 2, 5)
    eng.env_step(0.02, 40)
    eng.motion_validity(q1[:50], q2[:50], 6, 0.01, commit_edge=1)
    eng.self_collision_edges(q1[:300], q2[:300]); eng.check_edges(q1[:300], q2[:300]); eng.check_states(q[:300])
    t = eng.tree_create(); eng.tree_append(t, q[:3000]); eng.tree_nearest(q[3000:3100]); eng.tree_radius(t, q[5], 2.0, 5, capacity=64)
    eng.close()
# round 2: the large-batch path (coherent-warp sort, distance-field bootstrap, second pass) on a batch above both thresholds
rb = robots.xarm6()
mn, mx, g, tol = synthetic.sweep_obstacles(200, "D2", r_min=0.6)
eng = engine.Engine(rb); eng.set_obstacles(mn, mx, g, tol)
qq = synthetic.configs(rb, 70000)
eng.distance(qq, robots.DIST_PROFILE); eng.distance(qq, robots.DIST_MIN); eng.is_valid(qq)
e1, e2 = synthetic.edges(rb, 66000, 0.2)
eng.is_valid_edges(e1, e2)
eng.close()
rb, mn, mx, g, tol = scenes.xarm6_scenario1()
eng = engine.Engine(rb); eng.set_obstacles(mn, mx, g, tol)
cfg = planning.default_config(planning.RGBT_CONNECT, max_planning_time=5.0)
res, st = planning.solve_batch(eng, cfg, np.array([0, 0.7854, -2.3562, 0, 0, 0], np.float32), np.array([3.14159, 0.7854, -2.3562, 3.14159, 0, 3.14159], np.float32), seeds=np.arange(4))
out = planning.rrtx_grow(eng, planning.rrtx_config(eps_step=0.5, r_rewire=2.0, max_neighbors=50, device_index=1), np.array([3.14159, 0.7854, -2.3562, 3.14159, 0, 3.14159], np.float32),
                         synthetic.configs(rb, 150, seed=77))
print("sanitize probe done; planner success", [r["success"] for r in res], "rrtx states", out["stats"]["states"])
