#!/usr/bin/env python
"""bench.py — capsule-obstacle distance queries per second of the RealVectorSpaceCUDA engine (BASELINE.json metric).

One "step" = one pass of the hot path (RealVectorSpace::computeDistance with the per-link profile) over one batch of
2^20 synthetic xArm6 configurations against 1000 boxes (the BASELINE.json `configs[4]` sweep point).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (sm_100a kernels through the C ABI)
  python bench.py --impl reference ...                           the reference's CPU algorithm (oracle port) on the host cores
  torchrun --nproc-per-node N bench.py --gpus N ...              one process per GPU, configurations sharded, no collective
                                                                 on the data path (only the timing/pair-count reductions)

`value` counts the pair queries (distanceCapsuleToBox calls) the REFERENCE's own loop order would make for the batch
(link-major, early exit at the first collision, skipped ground pairs not counted) — counted by the kernel itself — divided by
the CUDA-event time of the K timed launches (max over ranks).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from rpmplv2_b200 import robots, synthetic  # noqa: E402

N_CONFIGS = 1 << 20
M_OBSTACLES = 1000
WORKLOAD = dict(workload="xarm6 sweep: 2^20 configs x 1000 AABBs (D2 boxes 5-30 cm, shell r>=0.9 m, + ground), computeDistance PROFILE",
                robot="xarm6", configs_per_step=N_CONFIGS, obstacles=M_OBSTACLES, obstacle_law="initRandomObstacles shell r_min=0.9 D2",
                mode="PROFILE", precision="REF32 (FK fp64, geometry fp32 no-FMA)", l2="flushed between timed steps (256 MiB write)",
                seeds=[hex(synthetic.SEED_CONFIGS), hex(synthetic.SEED_OBSTACLES)])
FLOP_PER_PAIR, FLOP_FK = 150.0, 630.0          # counting convention C1
# what the work-queue kernel executes per unit: 6 compares per scanned pair, ~25 flop per re-culled pair,
# ~60 per classified pair, 150 per segment-edge evaluation
FLOP_SCAN, FLOP_RECULL, FLOP_CLASSIFY, FLOP_EDGE = 6.0, 25.0, 60.0, 150.0
NCU_TRAFFIC_BYTES = 25_361_920 + 38_656        # dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full, profiles/r1i_*):
                                               # the configurations are read once; the 29 MB of results are still in the 126 MB L2 at kernel end


def scene():
    rb = robots.xarm6()
    return (rb,) + synthetic.sweep_obstacles(M_OBSTACLES, "D2", r_min=0.9)


class ClockSampler:
    """SM clock / throttle-reason sampler running DURING the timed region (B200_PROFILING.md clocks line).

    NVML in-process (one sample per ~2 ms: the timed region of this bench is only tens of milliseconds long, and an `nvidia-smi -lms`
    child needs longer than that to print its first line); falls back to the nvidia-smi child when pynvml is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.nvml, self.samples, self.run = index, [], None, None, [], False
        self.t0 = self.t1 = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml, self.handle = pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.run = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def _poll(self):
        nv, h = self.nvml, self.handle
        while self.run:
            try:
                self.samples.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetCurrentClocksThrottleReasons(h),
                                     nv.nvmlDeviceGetPowerUsage(h) / 1000.0))
            except Exception:
                pass
            time.sleep(0.002)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml:
            self.run = False
            self.thread.join(timeout=1.0)
            nv = self.nvml
            inside = [x for x in self.samples if self.t0 is not None and self.t1 is not None and self.t0 <= x[0] <= self.t1]
            if not inside:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
            sm = sorted(x[1] for x in inside)
            mask = 0
            for x in inside:
                mask |= int(x[2])
            return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)),
                    "reasons": [n for b, n in self.REASONS if mask & b], "samples": len(inside), "power_w_max": round(max(x[3] for x in inside), 2),
                    "source": "nvml, sampled inside the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 8 and r[1].isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(int(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[4 + k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(rows[0][2]), "reasons": reasons, "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows if r[3].replace(".", "").isdigit()), "source": "nvidia-smi -lms 20"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def cpu_baseline(threads: int, seconds_target: float = 12.0):
    """The oracle (CPU port of the reference algorithm) on a bounded sample of the same workload."""
    from oracle import oracle_binding as ob
    rb, mn, mx, g, tol = scene()
    o = ob.Oracle(rb, robots.default_params())
    o.set_obstacles(mn, mx, g, tol)
    probe = synthetic.configs(rb, 2048)
    t = time.perf_counter(); o.distance(probe, robots.DIST_PROFILE, threads=threads); rate = 2048 / (time.perf_counter() - t)
    n = int(max(4096, min(N_CONFIGS, rate * seconds_target)))
    q = synthetic.configs(rb, n)
    o.pair_count(True)
    t = time.perf_counter(); o.distance(q, robots.DIST_PROFILE, threads=threads); dt = time.perf_counter() - t
    pairs = o.pair_count(True)
    return {"value": pairs / dt, "unit": "pair-queries/s", "cores": threads, "kind": "port",
            "sample": f"first {n} of the {N_CONFIGS} configurations of one step, {threads} std::thread(s), {dt:.1f} s",
            "configs_per_s": n / dt}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port, all host threads) on bounded samples of the workload."""
    if rank != 0:
        return
    from oracle import oracle_binding as ob
    rb, mn, mx, g, tol = scene()
    threads = os.cpu_count() or 1
    o = ob.Oracle(rb, robots.default_params())
    o.set_obstacles(mn, mx, g, tol)
    probe = synthetic.configs(rb, 4096)
    t = time.perf_counter(); o.distance(probe, robots.DIST_PROFILE, threads=threads); rate = 4096 / (time.perf_counter() - t)
    budget = 150.0 / max(1, args.steps + args.warmup)
    n = int(max(4096, min(N_CONFIGS, rate * min(budget, 10.0))))
    total_pairs, total_t = 0, 0.0
    for step in range(args.warmup + args.steps):
        q = synthetic.configs(rb, n, first=step * N_CONFIGS)
        o.pair_count(True)
        t = time.perf_counter(); o.distance(q, robots.DIST_PROFILE, threads=threads); dt = time.perf_counter() - t
        if step >= args.warmup:
            total_pairs += o.pair_count(True); total_t += dt
    value = total_pairs / total_t
    sample = f"each step = first {n} of the step's {N_CONFIGS} configurations, {threads} host threads"
    print(json.dumps({"metric": "capsule_obstacle_distance_queries_per_s", "value": value, "unit": "pair-queries/s", "impl": "reference",
                      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": dict(WORKLOAD, cpu_sample=sample),
                      "cpu_baseline": {"value": value, "unit": "pair-queries/s", "cores": threads, "kind": "port", "sample": sample},
                      "e2e": {"value": value, "unit": "pair-queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="ref32", choices=["ref32", "fast32", "f64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cull", action="store_true", help="sequential one-thread-per-configuration kernel (no broad phase)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # libraries (NCCL's version banner, torch warnings) must not pollute stdout: the contract is ONE JSON line there.
    # Everything written to fd 1 during the run goes to stderr; the JSON line is written to the real stdout at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from rpmplv2_b200 import engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    prec = {"ref32": robots.PREC_REF32, "fast32": robots.PREC_FAST32, "f64": robots.PREC_F64}[args.precision]
    rb, mn, mx, g, tol = scene()
    eng = engine.Engine(rb, robots.default_params(prec), device=local)
    eng.set_obstacles(mn, mx, g, tol)
    eng.set_culling(not args.no_cull)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    n, L = rb.num_dofs, rb.num_links
    K, W = args.steps, args.warmup
    odt = torch.float64 if prec == robots.PREC_F64 else torch.float32

    # inputs resident in HBM before the timed region: one distinct batch per step, sharded by rank (weak scaling)
    d_q = torch.empty((K + W, N_CONFIGS, n), dtype=torch.float32, device=dev)
    for s in range(K + W):
        eng.generate_configs_dev(synthetic.SEED_CONFIGS, (rank * (K + W) + s) * N_CONFIGS, N_CONFIGS, d_q[s].data_ptr())
    d_dc = torch.empty(N_CONFIGS, dtype=odt, device=dev)
    d_prof = torch.empty((N_CONFIGS, L), dtype=odt, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def step(s):
        eng.distance_dev(d_q[s].data_ptr(), N_CONFIGS, robots.DIST_PROFILE, d_dc.data_ptr(), d_prof.data_ptr())

    peak32 = eng.measure_fma_peak(False)
    peak64 = eng.measure_fma_peak(True)
    for s in range(W):
        step(s)
    barrier()
    eng.work_stats(reset=True)
    launches0 = eng.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    sampler.mark_begin()
    for s in range(K):
        flush.fill_(s & 0xFF)                     # L2 flush, outside the event bracket
        ev[s][0].record(stream)
        step(W + s)
        ev[s][1].record(stream)
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    ms = [a.elapsed_time(b) for a, b in ev]
    t_rank = sum(ms) * 1e-3
    work = eng.work_stats(reset=True)
    pairs_rank = work["pair_queries"]
    launches = eng.launch_count() - launches0
    free_frac = float((d_dc > 0).float().mean().item())

    # end to end through the C ABI with HOST buffers (pinned): H2D of the step's configurations, kernel, D2H of d_c + profile
    h_q = torch.empty((N_CONFIGS, n), dtype=torch.float32).pin_memory()
    h_q.copy_(d_q[W].cpu())
    h_dc = torch.empty(N_CONFIGS, dtype=odt).pin_memory()
    h_prof = torch.empty((N_CONFIGS, L), dtype=odt).pin_memory()
    e2e_steps = max(3, min(K, 5))
    import ctypes
    lib = engine.load_library()
    fn = lib.rvs_distance_f64 if prec == robots.PREC_F64 else lib.rvs_distance
    fptr = ctypes.POINTER(ctypes.c_double if prec == robots.PREC_F64 else ctypes.c_float)

    def e2e_step():
        rc = fn(eng._h, ctypes.cast(h_q.data_ptr(), ctypes.POINTER(ctypes.c_float)), N_CONFIGS, robots.DIST_PROFILE,
                ctypes.cast(h_dc.data_ptr(), fptr), ctypes.cast(h_prof.data_ptr(), fptr), None)
        if rc != 0:
            raise RuntimeError(lib.rvs_last_error(eng._h))
    e2e_step()
    barrier()
    eng.pair_count(reset=True)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    t_e2e = time.perf_counter() - t0
    pairs_e2e = eng.pair_count(reset=True)
    esz = 8 if prec == robots.PREC_F64 else 4

    # supplementary (not part of the metric): isValid over the same batch, device-resident, median of 5 launches on this rank
    verdict = None
    if prec != robots.PREC_F64:
        d_v = torch.empty(N_CONFIGS, dtype=torch.uint8, device=dev)
        vms = []
        for s in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); eng.is_valid_dev(d_q[W + (s % K)].data_ptr(), N_CONFIGS, d_v.data_ptr()); b.record(stream)
            torch.cuda.synchronize()
            vms.append(a.elapsed_time(b))
        vmed = sorted(vms[1:])[2]
        verdict = {"ms_per_launch": vmed, "configs_per_s": N_CONFIGS / vmed * 1e3,
                   "collision_pair_queries_per_s": N_CONFIGS * synthetic.pair_queries(rb, g) / vmed * 1e3,
                   "valid_fraction": round(float(d_v.float().mean().item()), 4),
                   "kernel": "rvs::valid_wq_kernel (scan + pair / edge phases); nominal pairs L*M per configuration, no early-exit credit"}

    # reductions across ranks: time = max, pairs = sum
    red = torch.tensor([t_rank, t_e2e], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(pairs_rank), float(pairs_e2e), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    t_all, t_e2e_all = red.tolist()
    pairs_all, pairs_e2e_all, launches_all = tot.tolist()

    if rank == 0:
        peaks, peak_src = measured_peaks()
        flops_launch = (N_CONFIGS * FLOP_FK + pairs_rank / K * FLOP_PER_PAIR)
        avg_launch_s = t_rank / K
        achieved = flops_launch / avg_launch_s / 1e12
        peak = peak64 if prec == robots.PREC_F64 else peak32
        alg_bytes = N_CONFIGS * (4 * n + esz * (1 + L)) + M_OBSTACLES * 32
        if args.no_cull:
            kernel = "rvs::distance_kernel<%s,PROFILE> (one thread per configuration, every visited pair through the narrow phase)"
            executed = {"pairs_narrow_phase": pairs_rank / K, "flop": flops_launch, "tflops": achieved, "frac_of_peak": achieved / peak if peak else None}
        else:
            kernel = "rvs::distance_wq_kernel<%s,PROFILE,6> (broad-phase scan + warp work queues)"
            scan_tests = N_CONFIGS * (L * M_OBSTACLES - rb.ground_included)
            ex_flop = (N_CONFIGS * FLOP_FK + scan_tests * FLOP_SCAN + work["scan_survivors"] / K * FLOP_RECULL +
                       work["classified"] / K * FLOP_CLASSIFY + work["edges"] / K * FLOP_EDGE)
            executed = {"scan_tests": scan_tests, "scan_survivors": work["scan_survivors"] / K, "classified": work["classified"] / K,
                        "edges": work["edges"] / K, "flop": ex_flop, "tflops": ex_flop / avg_launch_s / 1e12,
                        "frac_of_peak": ex_flop / avg_launch_s / 1e12 / peak if peak else None}
        kernel = kernel % {"ref32": "double,float", "fast32": "float,float", "f64": "double,double"}[args.precision]
        out = {
            "metric": "capsule_obstacle_distance_queries_per_s", "value": pairs_all / t_all, "unit": "pair-queries/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": 1e3 * t_all / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"ref32": "f32", "fast32": "f32", "f64": "f64"}[args.precision], "data": "synthetic",
            "config": dict(WORKLOAD, precision=args.precision, collision_free_fraction=round(free_frac, 4),
                           pair_queries_per_config=round(pairs_rank / K / N_CONFIGS, 1), parallelism=f"configs sharded over {world} GPU(s)",
                           broad_phase="off" if args.no_cull else "on (results bit-identical to off)"),
            "configs_per_s": world * N_CONFIGS * K / t_all,
            "e2e": {"value": pairs_e2e_all / t_e2e_all, "unit": "pair-queries/s", "h2d_bytes_per_step": N_CONFIGS * 4 * n,
                    "d2h_bytes_per_step": N_CONFIGS * esz * (1 + L), "steps": e2e_steps, "ms_per_step": 1e3 * t_e2e_all / e2e_steps,
                    "api": "rvs_distance (host pointers, pinned), H2D + kernel + D2H inside the timed region"},
            "gpu_launches": int(launches_all),
            "is_valid": verdict,
            "roofline": {"bound": "fp32" if prec != robots.PREC_F64 else "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if peak else None, "traffic": NCU_TRAFFIC_BYTES,
                         "convention": "C1: 630 flop FK per config + 150 flop per reference pair query (SURVEY.md 8d)",
                         "peak_source": "FFMA/DFMA register microbenchmark measured in this run (rvs_measure_fma_peak)",
                         "kernel": kernel, "avg_launch_ms": 1e3 * avg_launch_s, "executed": executed,
                         "note": "C1 credits 150 flop to every pair query the reference would make; the broad phase answers ~97% of them with "
                                 "6 compares, so `frac` can exceed 1 - `executed` is what the kernel really issued (CUDA-core issue bound, not HBM)",
                         "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / avg_launch_s / 1e9,
                                 "peak_gbs": peaks.get("hbm_gbs"), "peak_source": peak_src}},
            "fma_peak_tflops": {"fp32": peak32, "fp64": peak64},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(os.cpu_count() or 1)
            out["cpu_baseline_1thread"] = cpu_baseline(1, seconds_target=6.0)
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
