#!/usr/bin/env python
"""bench_sweep.py — the S1 / S2 measurement matrix of SURVEY.md §8(d) on ONE GPU (supplementary to bench.py's headline line).

S1 "sweep": xArm6, N = 2^20 configurations (device-resident, stateless generator), M in {1, 10, 100, 1000} obstacles, size laws
D1 (1 cm cubes) / D2 (5-30 cm boxes), modes VERDICT (isValid), MIN, PROFILE; one 10^7 row; FULL at the xarm6 scenario size (M = 51).
S2 "planner-shaped": edge batches (E = 10^3, 10^5, EPS_STEP 0.1, 10 checks), spine sets (1 root x H targets, 5 layers) for xarm6 M = 51
and spatial_22dof M = 6, and the single-query latency of planar_2dof scenario_test through the host-pointer API.
Every row: CUDA-event time per call (median of `--reps`), reference-order pair queries per second where the notion applies.
Writes one JSON line per row to stdout and a markdown table to --out.
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

from rpmplv2_b200 import engine, robots, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.md"))
    ap.add_argument("--big", action="store_true", help="add the N = 10^7 row")
    args = ap.parse_args()
    import torch
    import scenes

    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    rows = []

    def timed(fn, reps=args.reps):
        fn(); torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(); b.record(stream); torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        return float(np.median(ms))

    def emit(**kw):
        rows.append(kw)
        print(json.dumps(kw), flush=True)

    # ---- S1 -------------------------------------------------------------------------------------------------------
    rb = robots.xarm6()
    n, L = rb.num_dofs, rb.num_links
    for prec_name, prec in (("ref32", robots.PREC_REF32), ("fast32", robots.PREC_FAST32)):
        eng = engine.Engine(rb, robots.default_params(prec), device=0)
        eng.set_stream(stream.cuda_stream)
        sizes = [1 << 20] + ([10_000_000] if args.big and prec_name == "ref32" else [])
        for N in sizes:
            d_q = torch.empty((N, n), dtype=torch.float32, device=dev)
            eng.generate_configs_dev(synthetic.SEED_CONFIGS, 0, N, d_q.data_ptr())
            d_dc = torch.empty(N, dtype=torch.float32, device=dev)
            d_prof = torch.empty((N, L), dtype=torch.float32, device=dev)
            d_v = torch.empty(N, dtype=torch.uint8, device=dev)
            for dist, r_min in (("D2", 0.9), ("D1", 0.0)):
                for M in (1, 10, 100, 1000):
                    if N > (1 << 20) and M != 1000:
                        continue
                    mn, mx, g, tol = synthetic.sweep_obstacles(M, dist, r_min=r_min)
                    eng.set_obstacles(mn, mx, g, tol)
                    for mode_name in ("VERDICT", "MIN", "PROFILE"):
                        if prec_name == "fast32" and (mode_name != "PROFILE" or dist != "D2"):
                            continue
                        eng.work_stats(reset=True)
                        if mode_name == "VERDICT":
                            fn = lambda: eng.is_valid_dev(d_q.data_ptr(), N, d_v.data_ptr())
                        else:
                            mode = robots.DIST_MIN if mode_name == "MIN" else robots.DIST_PROFILE
                            fn = lambda: eng.distance_dev(d_q.data_ptr(), N, mode, d_dc.data_ptr(), d_prof.data_ptr())
                        ms = timed(fn)
                        calls = args.reps + 1
                        pairs = eng.work_stats(reset=True)["pair_queries"] / calls if mode_name != "VERDICT" else None
                        nominal = N * synthetic.pair_queries(rb, g)
                        free = float((d_v > 0).float().mean().item()) if mode_name == "VERDICT" else float((d_dc > 0).float().mean().item())
                        emit(set="S1", robot="xarm6", precision=prec_name, N=N, M=M, law=f"{dist} r_min={r_min}", mode=mode_name, ms=ms,
                             configs_per_s=N / ms * 1e3, pair_queries_per_s=(pairs / ms * 1e3) if pairs else None,
                             nominal_pairs_per_s=nominal / ms * 1e3, collision_free=round(free, 4))
            del d_q, d_dc, d_prof, d_v
        # FULL at the xarm6 random-obstacle scenario size (ground + 50 cubes of 1 cm), HBM-write bound
        if prec_name == "ref32":
            N, M = 1 << 17, 51
            mn, mx, g, tol = synthetic.sweep_obstacles(M, "D1", r_min=0.3)
            eng.set_obstacles(mn, mx, g, tol)
            d_q = torch.empty((N, n), dtype=torch.float32, device=dev)
            eng.generate_configs_dev(synthetic.SEED_CONFIGS, 0, N, d_q.data_ptr())
            d_dc = torch.empty(N, dtype=torch.float32, device=dev); d_prof = torch.empty((N, L), dtype=torch.float32, device=dev)
            d_near = torch.empty((N, M, L, 6), dtype=torch.float32, device=dev)
            eng.work_stats(reset=True)
            ms = timed(lambda: eng.distance_dev(d_q.data_ptr(), N, robots.DIST_FULL, d_dc.data_ptr(), d_prof.data_ptr(), d_near.data_ptr()))
            pairs = eng.work_stats(reset=True)["pair_queries"] / (args.reps + 1)
            emit(set="S1", robot="xarm6", precision=prec_name, N=N, M=M, law="D1 r_min=0.3", mode="FULL", ms=ms, configs_per_s=N / ms * 1e3,
                 pair_queries_per_s=pairs / ms * 1e3, nominal_pairs_per_s=None, collision_free=round(float((d_dc > 0).float().mean().item()), 4),
                 hbm_write_gbs=N * (4 * (1 + L) + 24 * L * M) / ms / 1e6)
            del d_q, d_dc, d_prof, d_near
        eng.close()

    # ---- S2 (a) edge batches ------------------------------------------------------------------------------------------
    eng = engine.Engine(rb, robots.default_params(), device=0)
    eng.set_stream(stream.cuda_stream)
    mn, mx, g, tol = synthetic.sweep_obstacles(51, "D1", r_min=0.3)
    eng.set_obstacles(mn, mx, g, tol)
    for E in (1000, 100_000):
        q1, q2 = synthetic.edges(rb, E, 0.1)
        d1, d2 = torch.from_numpy(q1).to(dev), torch.from_numpy(q2).to(dev)
        d_v = torch.empty(E, dtype=torch.uint8, device=dev)
        ms = timed(lambda: eng.is_valid_edges_dev(d1.data_ptr(), d2.data_ptr(), E, d_v.data_ptr()))
        t0 = time.perf_counter(); eng.is_valid_edges(q1, q2); host_ms = (time.perf_counter() - t0) * 1e3
        emit(set="S2a", robot="xarm6", M=51, E=E, checks=10, ms=ms, host_call_ms=host_ms, edges_per_s=E / ms * 1e3, valid=round(float(d_v.float().mean().item()), 4))

    # ---- S2 (b) spine sets ----------------------------------------------------------------------------------------------
    for name, rbs, scene in (("xarm6", rb, (mn, mx, g, np.full_like(tol, 1000.0))), ("spatial_22dof", robots.spatial(22), None)):
        if scene is None:
            _, smn, smx, sg, stol = scenes.chain_scene(rbs, 6)
            scene = (smn, smx, sg, np.full_like(stol, 1000.0))
        e2 = engine.Engine(rbs, robots.default_params(), device=0)
        e2.set_obstacles(*scene)
        cand = synthetic.configs(rbs, 400)
        d, p, near = e2.distance(cand, robots.DIST_FULL)
        k = int(np.argmax(d))
        rng = np.random.default_rng(1)
        for H in (20, 70, 262):
            targets = (cand[k] + rng.normal(0, 1.5, (H, rbs.num_dofs))).astype(np.float32)
            t = []
            for _ in range(args.reps + 1):
                t0 = time.perf_counter()
                qr, dr, st = e2.extend_spines(cand[k], p[k], float(d[k]), near[k], targets)
                dd, _, _ = e2.distance(qr, robots.DIST_PROFILE)                      # the final real distance of every reached state
                t.append((time.perf_counter() - t0) * 1e3)
            emit(set="S2b", robot=name, M=len(scene[2]), H=H, layers=5, host_call_ms=float(np.median(t[1:])), spines_per_s=H / np.median(t[1:]) * 1e3,
                 reached=int((st == 3).sum()), advanced=int((st == 1).sum()), all_collision_free=bool((dd > 0).all()))
        e2.close()

    # ---- S2 (d) MotionValidity::check with the environment on the device (row f-2; config 4: spatial_22dof + 5 moving cubes) ----------
    from oracle import oracle_binding as ob      # CPU column only (checker / baseline, never the measured product path)
    for name, rbm in (("spatial_22dof", robots.spatial(22)), ("xarm6", rb)):
        reach = 1.0 if name == "xarm6" else float(np.abs(rbm.xyz).sum())
        sc = scenes.dynamic_scene(rbm, 5, 21, r_lo=0.45 * reach, r_hi=0.95 * reach, size=0.3 * reach)
        ep = engine.env_params("light_directions", ws_center=scenes.WS_CENTER, ws_radius=1.0, base_radius=0.3, robot_max_vel=1.5)
        em = engine.Engine(rbm, robots.default_params(), device=0)
        orc = ob.Oracle(rbm)
        K, dt = 10, 0.005
        for E in (1, 64, 4096):
            q1, q2 = synthetic.edges(rbm, E, length=0.3, seed=0x99)
            em.env_set(ep, **sc)
            orc.env_set(ep, **sc)
            t = []
            for _ in range(args.reps + 20):
                t0 = time.perf_counter(); v, f = em.motion_validity(q1, q2, K, dt); t.append((time.perf_counter() - t0) * 1e3)
            # what an integration WITHOUT the device environment pays: K x (host update + table upload + isValid launch)
            half = np.float32(0.5) * sc["dim"]
            t2 = []
            em2 = engine.Engine(rbm, robots.default_params(), device=0)
            em2.set_obstacles(sc["pos"] - half, sc["pos"] + half, sc["is_ground"], sc["min_dist_tol"])
            dist = np.linalg.norm(q2 - q1, axis=1, keepdims=True)
            for _ in range(args.reps + 5):
                t0 = time.perf_counter()
                pos = sc["pos"].copy()
                for k in range(1, K + 1):
                    pos = pos + np.float32(dt) * sc["vel"]                                   # stand-in for Environment::updateEnvironment on the host
                    em2.update_obstacle_aabbs(pos - half, pos + half)
                    em2.is_valid(q1 + (q2 - q1) * np.float32(k / K))
                t2.append((time.perf_counter() - t0) * 1e3)
            em2.close()
            t0 = time.perf_counter(); vo, fo = orc.motion_validity(q1, q2, K, dt); cpu_ms = (time.perf_counter() - t0) * 1e3
            emit(set="S2d", robot=name, M=6, E=E, checks=K, call="motion_validity (1 call)", host_call_ms=float(np.median(t[5:])),
                 per_point_path_ms=float(np.median(t2[2:])), cpu_oracle_1thread_ms=cpu_ms, edges_per_s=E / np.median(t[5:]) * 1e3,
                 valid=round(float(v.mean()), 4), identical_to_oracle=bool(np.array_equal(f, fo)))
        em.close()

    # ---- S2 (e) one DRGBT iteration's geometry (row f-1): computeDistance(q_current, true) + H generalized-bur spines + H real
    # distances (DRGBT.cpp:101, 272-320, 475-495) + the motion check with moving obstacles (MotionValidity.cpp:30-51);
    # spatial_22dof, ground + 5 moving cubes (BASELINE config 4); H = 10 + 2n - 2 = 52 and 10n = 220 (critical clearance)
    from oracle.oracle_space import OracleSpace      # CPU column only
    from rpmplv2_b200 import drgbt
    _horizon = synthetic.horizon_states
    rbd = robots.spatial(22)
    scd = scenes.dynamic_scene(rbd, 5, 21, r_lo=0.45 * 2.2, r_hi=0.95 * 2.2, size=0.3 * 2.2)
    scd["min_dist_tol"] = np.full(6, 1000.0, np.float32)
    epd = engine.env_params("light_directions", ws_center=scenes.WS_CENTER, ws_radius=2.2, base_radius=0.3, robot_max_vel=1.5)
    ed = engine.Engine(rbd, robots.default_params(), device=0)
    od = ob.Oracle(rbd)
    ed.env_set(epd, **scd)
    od.env_set(epd, **scd)
    candd = synthetic.configs(rbd, 300, seed=0xD6)
    q_cur = candd[int(np.argmax(od.distance(candd)[0]))]
    for H in (52, 220):
        hor = _horizon(rbd, q_cur, H, 11)
        tg, tc = [], []
        for _ in range(args.reps + 10):
            t0 = time.perf_counter()
            g = drgbt.generate_gbur(ed, q_cur, hor)
            v, f = ed.motion_validity(q_cur[None], g["q_reached"][:1], 10, 0.005)
            tg.append((time.perf_counter() - t0) * 1e3)
        for _ in range(3):
            t0 = time.perf_counter()
            go = drgbt.generate_gbur(OracleSpace(od), q_cur, hor)
            vo, fo = od.motion_validity(q_cur[None], go["q_reached"][:1], 10, 0.005)
            tc.append((time.perf_counter() - t0) * 1e3)
        emit(set="S2e", robot="spatial_22dof", M=6, H=H, layers=5, call="DRGBT iteration geometry (4 engine calls)", host_call_ms=float(np.median(tg[5:])),
             cpu_oracle_1thread_ms=float(np.median(tc)), spines_per_s=H / np.median(tg[5:]) * 1e3,
             same_statuses=bool(np.array_equal(g["spine_status"], go["spine_status"])), max_dq=float(np.abs(g["q_reached"] - go["q_reached"]).max()))
    ed.close()

    # ---- S2 (c) single-query latency, planar_2dof scenario_test ----------------------------------------------------------
    rb2, pmn, pmx, pg, ptol = scenes.planar2_scenario_test()
    e3 = engine.Engine(rb2, robots.default_params(), device=0)
    e3.set_obstacles(pmn, pmx, pg, ptol)
    q = synthetic.configs(rb2, 1)
    for label, fn in (("isValid(q)", lambda: e3.is_valid(q)), ("computeDistance(q) FULL", lambda: e3.distance(q, robots.DIST_FULL)),
                      ("isValid(q1,q2)", lambda: e3.is_valid_edges(q, q + np.float32(0.05)))):
        fn()
        t = []
        for _ in range(200):
            t0 = time.perf_counter(); fn(); t.append((time.perf_counter() - t0) * 1e6)
        emit(set="S2c", robot="planar_2dof", M=4, call=label, host_call_us_median=float(np.median(t)), host_call_us_p95=float(np.percentile(t, 95)))
    e3.close()
    eng.close()

    # ---- markdown ---------------------------------------------------------------------------------------------------------
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write("| set | robot | prec | N / E / H | M | law | mode / call | ms (device) | configs/s | pair-queries/s | free | note |\n|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            size = r.get("N") or r.get("E") or r.get("H") or 1
            note = ""
            if "host_call_ms" in r:
                note = f"host call {r['host_call_ms']:.3f} ms"
            if "host_call_us_median" in r:
                note = f"host call {r['host_call_us_median']:.1f} us (p95 {r['host_call_us_p95']:.1f})"
            if "hbm_write_gbs" in r:
                note = f"{r['hbm_write_gbs']:.0f} GB/s written"
            if r.get("set") == "S2e":
                note += f"; CPU oracle 1 thread {r['cpu_oracle_1thread_ms']:.3f} ms; same spine statuses {r['same_statuses']}"
            if "per_point_path_ms" in r:
                note += f"; K x (upload + isValid) {r['per_point_path_ms']:.3f} ms; CPU oracle 1 thread {r['cpu_oracle_1thread_ms']:.3f} ms; identical {r['identical_to_oracle']}"
            if "reached" in r:
                note += f"; reached {r['reached']}, advanced {r['advanced']}, collision-free {r['all_collision_free']}"
            pq = r.get("pair_queries_per_s") or r.get("nominal_pairs_per_s")
            f.write(f"| {r['set']} | {r['robot']} | {r.get('precision', 'ref32')} | {size} | {r.get('M', '')} | {r.get('law', '')} | {r.get('mode') or r.get('call', '')} | "
                    f"{r.get('ms', float('nan')):.3f} | {r.get('configs_per_s', r.get('edges_per_s', r.get('spines_per_s', float('nan')))):.3g} | "
                    f"{(pq if pq else float('nan')):.3g} | {r.get('collision_free', r.get('valid', ''))} | {note} |\n")


if __name__ == "__main__":
    main()
