#!/usr/bin/env python
"""bench_drgbt.py — BASELINE.json config 4: real-time replanning on spatial_22dof with dynamic random obstacles, timed per 50 ms iteration.

The loop has the SHAPE of planning::drbt::DRGBT::solve (/root/reference/src/planners/drbt/DRGBT.cpp:79-175) on the scene of
data/spatial_22dof/scenario_random_obstacles/scenario_random_obstacles.yaml (ground + random 0.3 m cubes at <= 0.3 m/s, motion_type
light_directions, WS_center (0, 0, 0.1), WS_radius 1.5) with configuration_drgbt.yaml's numbers (MAX_ITER_TIME 50 ms, INIT_HORIZON_SIZE 10
+ 2 n - 2 lateral states = 52 horizon states, D_CRIT 0.05, RESOLUTION_COLL_CHECK 0.01 -> ceil(0.3 * 0.05 / 0.01) = 2 validity points):
    every iteration
        computeDistance(q_current) + generateGBur over the horizon + computeDistance of every reached state   (drgbt.generate_gbur: 3 calls)
        the weight heuristic of computeNextState                                                               (drgbt.horizon_weights, host)
        a step of at most max_vel * MAX_ITER_TIME per joint towards the chosen reached state                   (host)
        MotionValidity::check(q_previous, q_current) while the obstacles move                                  (rvs_motion_validity: 1 call)
        when no horizon state is usable, or the predefined path is used up: replan with the static planner     (RGBT-Connect, 1 run)
DRGBT's own bookkeeping (horizon modification, bad / critical state handling, splines, the replanning heuristics) is NOT ported: this
times the state-space work one iteration asks for, through the calls a port of DRGBT onto RealVectorSpaceCUDA would make.
Two arms run the same loop with the same seeds: the CUDA engine, and the CPU oracle (the reference's RealVectorSpace algorithm, one
call at a time on one host thread — the reference's execution model).  Reported: time per iteration (median / p95 / max), its split,
the share of iterations inside the 50 ms budget, and the replanning calls.  A reported baseline next to bench.py, not the target.

  python bench_drgbt.py [--iters 300] [--obstacles 5] [--no-cpu]
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from rpmplv2_b200 import drgbt, planning, robots, synthetic  # noqa: E402
from rpmplv2_b200.engine import env_params  # noqa: E402

MAX_ITER_TIME, INIT_HORIZON_SIZE, RESOLUTION_COLL_CHECK = 0.050, 10, 0.01      # configuration_drgbt.yaml
WS_CENTER, WS_RADIUS, OBS_DIM, OBS_MAX_VEL, ROBOT_MAX_VEL = (0.0, 0.0, 0.1), 1.5, 0.3, 0.3, 3.1415
Q_START = np.array([0, 1.57] + [0] * 20, np.float32)
Q_GOAL = np.array([3.14, 1.57] + [0] * 20, np.float32)


def scene(M: int, seed: int):
    """Ground box + M random cubes (apps/CommonFunctions.h:115-170: position r * (cos fi sin th, sin fi sin th, cos th) around WS_center,
    velocity in a random direction with a random share of max_vel), kept clear of the start and goal poses."""
    rng = np.random.default_rng(seed)
    pos, vel = [[0.0, 0.0, -0.05]], [[0.0, 0.0, 0.0]]
    while len(pos) < M + 1:
        r, fi, th = rng.uniform(0.45, 1.0) * WS_RADIUS, rng.uniform(0, 2 * math.pi), rng.uniform(0, 0.5 * math.pi)
        p = np.array(WS_CENTER) + r * np.array([math.cos(fi) * math.sin(th), math.sin(fi) * math.sin(th), math.cos(th)])
        if p[2] < 0.25 or abs(p[0]) < 0.35 and abs(p[1]) < 0.35:               # not on the ground plate, not above the base column
            continue
        v = rng.normal(size=3)
        pos.append(list(p)); vel.append(list(v / np.linalg.norm(v) * rng.uniform(0.3, 1.0) * OBS_MAX_VEL))
    dim = np.array([[1.5, 1.5, 0.1]] + [[OBS_DIM] * 3] * M, np.float32)
    return dict(dim=dim, pos=np.array(pos, np.float32), vel=np.array(vel, np.float32), max_vel=np.array([0.0] + [OBS_MAX_VEL] * M, np.float32),
                is_dynamic=np.array([0] + [1] * M, np.uint8), is_ground=np.array([1] + [0] * M, np.uint8),
                min_dist_tol=np.array([0.05] + [1000.0] * M, np.float32))


class Arm:
    """One side of the comparison: `space` answers generate_gbur's calls, `mv` the motion check, `plan` the static planner."""

    def __init__(self, name, space, mv, plan, env_get):
        self.name, self.space, self.mv, self.plan, self.env_get = name, space, mv, plan, env_get


def run_loop(arm: Arm, rb, iters: int, seed: int):
    n = rb.num_dofs
    K = int(math.ceil(OBS_MAX_VEL * MAX_ITER_TIME / RESOLUTION_COLL_CHECK))
    sp = drgbt.default_spine_params()
    q_cur = Q_START.copy()
    path, path_i = np.zeros((0, n), np.float32), 0
    d_prev = np.zeros(0, np.float32)
    d_max_mean, collisions, replans, goals_reached = 0.0, 0, 0, 0
    q_goal = Q_GOAL.copy()
    t_iter, t_gbur, t_mv, t_plan = [], [], [], []
    for it in range(iters):
        t0 = time.perf_counter()
        tp = 0.0
        if path_i >= len(path):                                                                   # replan(): the static planner from q_current
            t = time.perf_counter()
            res = arm.plan(q_cur, q_goal, seed + it)
            tp = time.perf_counter() - t
            replans += 1
            path, path_i = (res["path"][1:], 0) if res["success"] and res["path"] is not None and len(res["path"]) > 1 else (np.zeros((0, n), np.float32), 0)
        ahead = path[path_i:path_i + INIT_HORIZON_SIZE]
        lateral = synthetic.horizon_states(rb, q_cur, 2 * n - 2 + INIT_HORIZON_SIZE - len(ahead), seed * 7919 + it)
        horizon = np.vstack([ahead, lateral]).astype(np.float32)
        index = np.concatenate([np.arange(path_i, path_i + len(ahead)), np.full(len(lateral), -1)])
        t = time.perf_counter()
        g = drgbt.generate_gbur(arm.space, q_cur, horizon, sp)
        tg = time.perf_counter() - t
        d_previous = d_prev if len(d_prev) == len(horizon) else np.zeros(len(horizon), np.float32)
        w, best, d_max_mean = drgbt.horizon_weights(g["q_reached"], g["d_c"], d_previous, index, g["status"], q_goal, d_max_mean, it)
        d_prev = g["d_c"]
        q_prev = q_cur.copy()
        if best >= 0:
            step = g["q_reached"][best] - q_cur
            lim = np.float32(ROBOT_MAX_VEL * MAX_ITER_TIME)
            scale = min(1.0, float(lim / max(np.abs(step).max(), 1e-9)))
            q_cur = (q_cur + np.float32(scale) * step).astype(np.float32)
            if index[best] >= 0 and scale == 1.0 and bool(g["is_reached"][best]):
                path_i = int(index[best]) + 1                                                     # that path node is behind us now
        else:
            path_i = len(path)                                                                    # nothing usable: replan next iteration
        t = time.perf_counter()
        valid, _ = arm.mv(q_prev[None], q_cur[None], K, MAX_ITER_TIME / K)
        tm = time.perf_counter() - t
        collisions += int(not valid[0])
        t_iter.append(time.perf_counter() - t0); t_gbur.append(tg); t_mv.append(tm); t_plan.append(tp)
        if float(np.abs(q_cur - q_goal).max()) < 1e-3:                                            # the next test of the scenario file: back again
            goals_reached += 1
            q_goal = (Q_START if goals_reached % 2 else Q_GOAL).copy()
            path, path_i = np.zeros((0, n), np.float32), 0
    ms = lambda a: 1e3 * np.asarray(a)
    ti = ms(t_iter)
    return dict(arm=arm.name, iterations=len(ti), goals_reached=goals_reached, collisions=collisions, replans=replans, validity_points=K,
                horizon_states=2 * n - 2 + INIT_HORIZON_SIZE,
                iter_ms=dict(median=float(np.median(ti)), mean=float(ti.mean()), p95=float(np.percentile(ti, 95)), max=float(ti.max())),
                within_50ms=float((ti <= 1e3 * MAX_ITER_TIME).mean()),
                no_replan_iter_ms_median=float(np.median(ti[ms(t_plan) == 0])) if (ms(t_plan) == 0).any() else None,
                split_ms_mean=dict(generate_gbur=float(ms(t_gbur).mean()), motion_validity=float(ms(t_mv).mean()),
                                   replanning=float(ms(t_plan).sum() / len(ti)), replanning_per_call=float(ms(t_plan)[ms(t_plan) > 0].mean()) if replans else 0.0),
                final_obstacle_checksum=float(np.asarray(arm.env_get()[0], np.float64).sum()))


def run(iters: int = 300, obstacles: int = 5, seed: int = 4, cpu: bool = True, gpu: bool = True, device: int = 0) -> dict:
    rb = robots.spatial(22)
    sc = scene(obstacles, seed)
    ep = env_params("light_directions", ws_center=WS_CENTER, ws_radius=WS_RADIUS, base_radius=0.3, robot_max_vel=ROBOT_MAX_VEL)
    cfg = planning.default_config(planning.RGBT_CONNECT, max_planning_time=MAX_ITER_TIME * 20)
    rows = []
    if gpu:
        from rpmplv2_b200.engine import Engine
        eng = Engine(rb, device=device)
        eng.env_set(ep, **sc)
        arm = Arm("RealVectorSpaceCUDA (B200)", eng, lambda a, b, K, dt: eng.motion_validity(a, b, K, dt, commit_edge=0),
                  lambda qs, qg, s: planning.solve_batch(eng, cfg, qs, qg, seeds=[s])[0][0], eng.env_get)
        run_loop(arm, rb, 20, seed)                                                               # warm-up: allocations, module load
        eng.env_set(ep, **sc)
        launches = eng.launch_count()
        row = run_loop(arm, rb, iters, seed)
        row["kernel_launches_per_iteration"] = (eng.launch_count() - launches) / row["iterations"]
        rows.append(row)
        eng.close()
    if cpu:
        from oracle import oracle_binding as ob                                                   # CPU baseline leg only
        from oracle.oracle_space import OracleSpace
        o = ob.Oracle(rb, robots.default_params())
        o.env_set(ep, **sc)
        arm = Arm("RealVectorSpace (CPU oracle port, 1 thread)", OracleSpace(o), lambda a, b, K, dt: o.motion_validity(a, b, K, dt, commit_edge=0),
                  lambda qs, qg, s: ob.plan_solve_batch(o, cfg, qs, qg, seeds=[s])[0][0], o.env_get)
        rows.append(run_loop(arm, rb, iters, seed))
    out = dict(bench="drgbt_iteration", config="spatial_22dof scenario_random_obstacles, light_directions, %d obstacles + ground" % obstacles,
               max_iter_time_ms=1e3 * MAX_ITER_TIME, arms=rows)
    if len(rows) == 2:
        out["iteration_speedup_median"] = rows[1]["iter_ms"]["median"] / rows[0]["iter_ms"]["median"]
        out["same_run"] = rows[0]["final_obstacle_checksum"] == rows[1]["final_obstacle_checksum"] and rows[0]["iterations"] == rows[1]["iterations"]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--obstacles", type=int, default=5, help="random_obstacles.init_num of the scenario file is 5; larger values stress the engine")
    ap.add_argument("--seed", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gpu", action="store_true")
    args = ap.parse_args()
    print(json.dumps(run(args.iters, args.obstacles, args.seed, cpu=not args.no_cpu, gpu=not args.no_gpu)))


if __name__ == "__main__":
    main()
