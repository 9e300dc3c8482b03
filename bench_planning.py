#!/usr/bin/env python
"""bench_planning.py — end-to-end planning time of the batched planners on the CUDA state space, next to the same planner code
driven by the CPU oracle (the reference's RealVectorSpace algorithm) on the host cores (BASELINE.json: "planning time vs CPU").

  python bench_planning.py [--runs P] [--json]

For every scenario P independent runs (seeds 0..P-1) are planned
  gpu_lockstep : all P runs in ONE pool on the GPU (rvsplan_solve_batch)          -> wall time of the call
  gpu_single   : run 0..min(P,8)-1 one at a time on the GPU                       -> mean wall time per run (latency)
  cpu_1thread  : runs one at a time on ONE host thread (the reference's execution model) -> mean time per run
  cpu_all      : the P runs spread over all host threads                          -> wall time of the call
This is a reported baseline next to the kernel numbers of bench.py, not the optimisation target.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from rpmplv2_b200 import planning, robots, synthetic  # noqa: E402


def _pillars():
    dim = np.array([[1.5, 1.5, 0.1]] + [[0.5, 0.5, 1.0]] * 4, np.float32)
    pos = np.array([[0, 0, -0.05], [0.55, 0.55, 0.5], [-0.55, 0.55, 0.5], [0.55, -0.55, 0.5], [-0.55, -0.55, 0.5]], np.float32)
    mn, mx = synthetic.box_aabbs(dim, pos)
    return mn, mx, np.array([1, 0, 0, 0, 0], np.uint8), np.array([0.05, 1000, 1000, 1000, 1000], np.float32)


def scenarios():
    """BASELINE.json configs[0..2] (geometry of the reference's data/*/scenario*.yaml) plus a cluttered xArm6 scene."""
    out = []
    dim = np.tile(np.array([[0.5, 0.5, 0.1]], np.float32), (4, 1))
    pos = np.array([[1, 1, 0], [-1, 1, 0], [1, -1, 0], [-1, -1, 0]], np.float32)
    mn, mx = synthetic.box_aabbs(dim, pos)
    out.append(dict(name="planar_2dof scenario_test", planner=planning.RGBT_CONNECT, robot=robots.planar(2),
                    obstacles=(mn, mx, np.zeros(4, np.uint8), np.full(4, 1000.0, np.float32)), q_start=[-2, -2.5], q_goal=[2, 2.5]))
    out.append(dict(name="xarm6 scenario1 (ground + 4 pillars)", planner=planning.RGBT_CONNECT, robot=robots.xarm6(), obstacles=_pillars(),
                    q_start=[0, 0.7854, -2.3562, 0, 0, 0], q_goal=[3.14159, 0.7854, -2.3562, 3.14159, 0, 3.14159]))
    out.append(dict(name="spatial_10dof scenario1", planner=planning.RBT_CONNECT, robot=robots.spatial(10), obstacles=_pillars(),
                    q_start=[0, 1.57, 0, 0, 0, 0, 0, 0, 0, 0], q_goal=[3.14, 1.57, 0, 0, 0, 0, 0, 0, 0, 0]))
    for M in (51, 1000):
        mn, mx, g, tol = synthetic.sweep_obstacles(M, "D1" if M == 51 else "D2", r_min=0.0 if M == 51 else 0.55)
        out.append(dict(name=f"xarm6 random obstacles M={M}", planner=planning.RGBT_CONNECT, robot=robots.xarm6(), obstacles=(mn, mx, g, tol),
                        q_start=[0, 0.7854, -2.3562, 0, 0, 0], q_goal=[3.14159, 0.7854, -2.3562, 3.14159, 0, 3.14159], prune_ends=True))   # scenario1's ends: free of self-collision
    return out


def _prune(eng_or_orc, sc):
    """drop random obstacles that collide with q_start / q_goal (the reference re-draws those, apps/CommonFunctions.h:102-110)"""
    mn, mx, g, tol = sc["obstacles"]
    ends = np.asarray([sc["q_start"], sc["q_goal"]], np.float32)
    keep = np.ones(len(mn), bool)
    for j in range(len(mn)):
        if g[j]:
            continue
        eng_or_orc.set_obstacles(mn[j:j + 1], mx[j:j + 1], g[j:j + 1], tol[j:j + 1])
        if not eng_or_orc.is_valid(ends).all():
            keep[j] = False
    sc["obstacles"] = (mn[keep], mx[keep], g[keep], tol[keep])


def run(P: int = 64, device: int = 0, cpu: bool = True, max_time: float = 30.0, only=None, skip=None):
    from rpmplv2_b200 import engine
    threads = os.cpu_count() or 1
    rows = []
    for sc in scenarios():
        if (only and only not in sc["name"]) or (skip and skip in sc["name"]):
            continue
        rb = sc["robot"]
        eng = engine.Engine(rb, robots.default_params(), device=device)
        if sc.get("prune_ends"):
            _prune(eng, sc)
        mn, mx, g, tol = sc["obstacles"]
        eng.set_obstacles(mn, mx, g, tol)
        cfg = planning.default_config(sc["planner"], max_planning_time=max_time)
        seeds = np.arange(P)
        planning.solve_batch(eng, cfg, sc["q_start"], sc["q_goal"], seeds=seeds[:2])            # warm-up (allocations, module load)
        t = time.perf_counter()
        res, st = planning.solve_batch(eng, cfg, sc["q_start"], sc["q_goal"], seeds=seeds)
        t_lock = time.perf_counter() - t
        k = min(P, 8)
        t = time.perf_counter()
        single = [planning.solve_batch(eng, cfg, sc["q_start"], sc["q_goal"], seeds=seeds[i:i + 1])[0][0] for i in range(k)]
        t_single = (time.perf_counter() - t) / k
        row = dict(scenario=sc["name"], planner=planning.KIND_NAMES[sc["planner"]], robot=rb.name, obstacles=int(len(mn)), runs=P,
                   success=int(sum(r["success"] for r in res)), mean_iterations=float(np.mean([r["iterations"] for r in res])),
                   mean_states=float(np.mean([r["states"] for r in res])),
                   queries_per_run=dict(distance=float(np.mean([r["num_distances"] for r in res])), spines=float(np.mean([r["num_spines"] for r in res])),
                                        edges=float(np.mean([r["num_edges"] for r in res]))),
                   gpu_lockstep=dict(wall_s=t_lock, per_run_ms=1e3 * t_lock / P, rounds=st["rounds"], backend_calls=st["backend_calls"],
                                     kernel_launches=st["kernel_launches"], us_per_backend_call=1e6 * t_lock / max(1, st["backend_calls"])),
                   gpu_single=dict(per_run_ms=1e3 * t_single, runs=k, success=int(sum(r["success"] for r in single))))
        # a pool that ran into max_planning_time reports the budget, not a speed: compare `success` within the budget instead
        row["gpu_lockstep"]["time_limited"] = bool(t_lock >= 0.98 * max_time)
        row["planning_time_budget_s"] = max_time
        if cpu:
            from oracle import oracle_binding as ob
            o = ob.Oracle(rb, robots.default_params())
            o.set_obstacles(mn, mx, g, tol)
            k1 = min(P, 8)
            t = time.perf_counter()
            r1, _ = ob.plan_solve_batch(o, cfg, sc["q_start"], sc["q_goal"], seeds=seeds[:k1], threads=1)
            t1 = (time.perf_counter() - t) / k1
            t = time.perf_counter()
            ra, _ = ob.plan_solve_batch(o, cfg, sc["q_start"], sc["q_goal"], seeds=seeds, threads=threads)
            ta = time.perf_counter() - t
            same = sum(1 for a, b in zip(res, ra) if a["iterations"] == b["iterations"] and a["states"] == b["states"])
            row["cpu_1thread"] = dict(per_run_ms=1e3 * t1, runs=k1, success=int(sum(r["success"] for r in r1)), kind="port (oracle)")
            row["cpu_all"] = dict(wall_s=ta, per_run_ms=1e3 * ta / P, cores=threads, success=int(sum(r["success"] for r in ra)), kind="port (oracle)")
            row["identical_runs_gpu_vs_cpu"] = same
            row["speedup_lockstep_vs_cpu_all"] = ta / t_lock
            row["speedup_single_vs_cpu_1thread"] = t1 / t_single
        rows.append(row)
        eng.close()
    return rows


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", type=int, default=64)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--only", default=None)
    ap.add_argument("--skip", default=None, help="skip scenarios whose name contains this (e.g. 'M=1000' for large --runs)")
    args = ap.parse_args()
    for row in run(args.runs, cpu=not args.no_cpu, only=args.only, skip=args.skip):
        print(json.dumps(row))
