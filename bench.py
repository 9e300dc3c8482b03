#!/usr/bin/env python
"""bench.py — capsule-obstacle distance queries per second of the RealVectorSpaceCUDA engine (BASELINE.json metric).

One "step" = one pass of the hot path (RealVectorSpace::computeDistance with the per-link profile, RealVectorSpace.cpp:318-392) over
one batch of 10^7 synthetic xArm6 configurations against 1000 boxes (BASELINE.json `configs[4]`, the batched sweep).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (sm_100a kernels through the C ABI), weak scaling
  python bench.py --scaling strong ...                           ONE fixed 10^8-configuration batch sharded over the ranks, results
                                                                 gathered to rank 0 over NVLink inside the timed region
  python bench.py --impl reference ...                           the reference's own CPU RealVectorSpace::computeDistance
                                                                 (oracle/_ref/librvs_ref.so, compiled from /root/reference at -O3; the
                                                                 oracle port when that library is absent) on all host threads
  torchrun --nproc-per-node N bench.py --gpus N ...              one process per GPU, configurations sharded, no collective on the
                                                                 data path (only the timing / pair-count reductions)

`value` counts the pair queries (distanceCapsuleToBox calls) the REFERENCE's own loop order makes for the batch (link-major, early
exit at the first collision, skipped ground pairs not counted) — counted by the kernel itself — divided by the CUDA-event time of
the K timed launches (max over ranks).  The JSON line also carries:
  roofline       hardware fraction of the dominant kernel: warp instructions issued per second against the issue peak of the device
                 (the kernel is compare / select bound: ALU-pipe issue, not HBM and not FMA), instruction count and DRAM traffic taken
                 from the `ncu --set full` capture of THIS tree under profiles/r2/, time measured live; the C1 "credited" figure and the
                 real FP32 arithmetic rate are given beside it
  workloads      the S1-as-written sweep points (obstacle centres from r = 0) and a near-field collision-free point
  parity_sample  2^14 configurations of the last timed step re-computed by the CPU oracle
  cpu_baseline   the reference's own code (kind "reference") and the oracle port on the host cores
"""
from __future__ import annotations

import argparse
import ctypes
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
from rpmplv2_b200.sharding import shard_range  # noqa: E402

N_CONFIGS = int(os.environ.get("RVS_BENCH_CONFIGS", 10_000_000))           # configurations per step and rank (weak scaling)
N_STRONG = int(os.environ.get("RVS_BENCH_STRONG_CONFIGS", 100_000_000))    # ONE batch over all ranks (--scaling strong)
N_FULL = int(os.environ.get("RVS_BENCH_FULL_CONFIGS", 1 << 20))                # configurations per launch of the `full_mode` point (7.7 GB of nearest points)
N_EXTRA = int(os.environ.get("RVS_BENCH_EXTRA_CONFIGS", 1 << 22))          # configurations per launch of the `workloads` points
MAX_RESIDENT_BATCHES = 24
FLOP_PER_PAIR, FLOP_FK = 150.0, 630.0          # counting convention C1
NCU_PROFILE = os.path.join(ROOT, "profiles", "r2", "ncu_distance_wq_headline.json")

# name -> obstacle law.  `headline` is the configuration the metric is quoted on (round-1 continuity: obstacles in the outer shell of the
# workspace, so that 18 % of the configurations are collision-free and every one of their 5998 pairs counts); the s1_* points are
# SURVEY.md §8d S1 exactly as written (centres r = u*R from 0: the arm is buried, nearly every configuration collides early);
# near_field_free keeps only collision-free configurations among obstacles inside the arm's reach (no early exit, no far-field).
SCENES = {
    "headline": dict(M=1000, dist="D2", r_min=0.9, r_max=None, free_only=False),
    "s1_m1000_d1": dict(M=1000, dist="D1", r_min=0.0, r_max=None, free_only=False),
    "s1_m1000_d2": dict(M=1000, dist="D2", r_min=0.0, r_max=None, free_only=False),
    "s1_m100_d1": dict(M=100, dist="D1", r_min=0.0, r_max=None, free_only=False),
    "s1_m100_d2": dict(M=100, dist="D2", r_min=0.0, r_max=None, free_only=False),
    "near_field_free": dict(M=100, dist="D1", r_min=0.0, r_max=0.8, free_only=True),
}
M_OBSTACLES = SCENES["headline"]["M"]


def workload_config(n_gpus: int, scaling: str):
    """The `config` object: identical for both arms (the driver compares them)."""
    per = N_CONFIGS if scaling == "weak" else N_STRONG
    return dict(workload=f"xarm6 sweep: {per:.0e} configs x 1000 AABBs (D2 boxes 5-30 cm, shell r>=0.9 m, + ground), computeDistance PROFILE".replace("e+0", "e"),
                robot="xarm6", configs_per_step=per, configs_per_step_is="per GPU" if scaling == "weak" else "whole job (sharded)",
                obstacles=M_OBSTACLES, obstacle_law="initRandomObstacles shell r_min=0.9 D2", mode="PROFILE",
                arithmetic="FK fp64, geometry fp32 without FMA contraction (the reference's x86-64 build)",
                l2="inputs of one step (240 MB) exceed the 126 MB L2 and every step reads a different batch: no flush needed",
                seeds=[hex(synthetic.SEED_CONFIGS), hex(synthetic.SEED_OBSTACLES)], parallelism=f"configs sharded over {n_gpus} GPU(s)",
                scaling=scaling)


def scene(name: str = "headline"):
    s = SCENES[name]
    rb = robots.xarm6()
    return (rb,) + synthetic.sweep_obstacles(s["M"], s["dist"], r_min=s["r_min"], r_max=s["r_max"])


class ClockSampler:
    """SM clock / throttle-reason sampler running DURING the timed region (B200_PROFILING.md clocks line).  NVML in-process (one sample
    per ~2 ms); falls back to an nvidia-smi child when pynvml is missing."""
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
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


# ---- CPU arms ------------------------------------------------------------------------------------------------------------------------
f32p, u8p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint8)
REF_SO = os.path.join(ROOT, "oracle", "_ref", "librvs_ref.so")


class CpuArm:
    """RealVectorSpace::computeDistance (PROFILE outputs) on host threads: the reference's own code compiled from /root/reference
    (oracle/_ref/librvs_ref.so, kind "reference") when that library travelled with the tree, else the oracle port (kind "port").
    The pair-query count of a sample always comes from the oracle's counter (identical outputs => identical loop trips)."""

    def __init__(self, prefer_reference: bool = True):
        from oracle import oracle_binding as ob
        self.rb, self.mn, self.mx, self.g, self.tol = scene()
        self.o = ob.Oracle(self.rb, robots.default_params())
        self.o.set_obstacles(self.mn, self.mx, self.g, self.tol)
        self.ref = None
        if prefer_reference and os.path.exists(REF_SO):
            lib = ctypes.CDLL(REF_SO)
            if hasattr(lib, "rvsref_distance_mt"):
                self.ref = lib
        self.kind = "reference" if self.ref is not None else "port"
        self.what = ("base::RealVectorSpace::computeDistance compiled from /root/reference at -O3 without FMA contraction (oracle/_ref/librvs_ref.so; "
                     "FK by the harness robot)" if self.ref is not None else "oracle port of RealVectorSpace::computeDistance (oracle/rvs_oracle.hpp), g++ -O3 without FMA contraction")

    def pairs(self, q):
        self.o.pair_count(True)
        self.o.distance(q, robots.DIST_PROFILE, threads=os.cpu_count() or 1)
        return self.o.pair_count(True)

    def run(self, q, threads: int, use_port: bool = False):
        """-> seconds for the batch q"""
        if self.ref is None or use_port:
            t = time.perf_counter(); self.o.distance(q, robots.DIST_PROFILE, threads=threads); return time.perf_counter() - t
        n, N = self.rb.num_dofs, len(q)
        d, p = np.empty(N, np.float32), np.empty((N, n), np.float32)
        desc = self.rb.to_desc()
        P = lambda a, t: np.ascontiguousarray(a).ctypes.data_as(t)
        q = np.ascontiguousarray(q, np.float32)
        t = time.perf_counter()
        rc = self.ref.rvsref_distance_mt(ctypes.byref(desc), ctypes.c_float(1e-4), ctypes.c_float(1000.0), P(self.mn, f32p), P(self.mx, f32p), P(self.g, u8p),
                                         P(self.tol, f32p), len(self.g), q.ctypes.data_as(f32p), ctypes.c_int64(N), threads, d.ctypes.data_as(f32p),
                                         p.ctypes.data_as(f32p))
        dt = time.perf_counter() - t
        if rc != 0:
            raise RuntimeError("rvsref_distance_mt failed")
        return dt

    def sample(self, threads: int, seconds_target: float, use_port: bool = False, first: int = 0):
        probe = synthetic.configs(self.rb, 256 * threads, first=first)
        rate = len(probe) / self.run(probe, threads, use_port)
        n = int(max(1024, min(N_CONFIGS, rate * seconds_target)))
        q = synthetic.configs(self.rb, n, first=first)
        dt = self.run(q, threads, use_port)
        return n, dt, self.pairs(q)


def cpu_baseline(seconds_target: float = 10.0):
    arm = CpuArm()
    cores = os.cpu_count() or 1
    out = {}
    n, dt, pairs = arm.sample(cores, seconds_target)
    out["cpu_baseline"] = {"value": pairs / dt, "unit": "pair-queries/s", "cores": cores, "kind": arm.kind, "what": arm.what,
                           "sample": f"first {n} of the {N_CONFIGS} configurations of one step, {cores} std::thread(s), {dt:.1f} s", "configs_per_s": n / dt}
    n, dt, pairs = arm.sample(1, 5.0)
    out["cpu_baseline_1thread"] = {"value": pairs / dt, "unit": "pair-queries/s", "cores": 1, "kind": arm.kind,
                                   "sample": f"first {n} configurations, 1 thread (the reference's own execution model), {dt:.1f} s", "configs_per_s": n / dt}
    if arm.kind == "reference":
        n, dt, pairs = arm.sample(cores, 6.0, use_port=True)
        out["cpu_baseline_port"] = {"value": pairs / dt, "unit": "pair-queries/s", "cores": cores, "kind": "port",
                                    "sample": f"first {n} configurations, {cores} std::thread(s), {dt:.1f} s: the oracle restatement (no per-pair heap allocations)",
                                    "configs_per_s": n / dt}
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU path on all host threads, every step a bounded sample of the step's configurations."""
    if rank != 0:
        return
    arm = CpuArm()
    threads = os.cpu_count() or 1
    steps = args.warmup + args.steps
    per_step = min(150.0 / max(1, steps), 10.0)
    probe = synthetic.configs(arm.rb, 256 * threads)
    rate = len(probe) / arm.run(probe, threads)
    n = int(max(1024, min(N_CONFIGS, rate * per_step)))
    total_pairs, total_t = 0, 0.0
    for step in range(steps):
        q = synthetic.configs(arm.rb, n, first=step * N_CONFIGS)
        dt = arm.run(q, threads)
        if step >= args.warmup:
            total_pairs += arm.pairs(q); total_t += dt
    value = total_pairs / total_t
    sample = f"each step = first {n} of the step's {N_CONFIGS} configurations, {threads} host threads, {arm.what}"
    print(json.dumps({"metric": "capsule_obstacle_distance_queries_per_s", "value": value, "unit": "pair-queries/s", "impl": "reference",
                      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
                      "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": workload_config(args.gpus, args.scaling),
                      "cpu_baseline": {"value": value, "unit": "pair-queries/s", "cores": threads, "kind": arm.kind, "sample": sample},
                      "e2e": {"value": value, "unit": "pair-queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


# ---- BASELINE.json's second metric: planning time next to the CPU arm, same run -------------------------------------------------------
def planning_summary(device: int):
    """`planning time vs CPU` for BASELINE configs 1-4, bounded to ~1 minute: 64 planner runs per shipped scene (bench_planning.py: all 64 in one
    lockstep pool on the GPU, single runs on the GPU, single runs on ONE host thread — the reference's execution model —, the 64 runs over all
    host threads; the CPU arm is the same planner code over the oracle port: cpu_baseline leg) and the DRGBT-shaped replanning loop of
    bench_drgbt.py per 50 ms iteration.  Config 2's RGBMT* and DRGBT's own bookkeeping are not restated: RGBT-Connect / the iteration's
    state-space calls stand in.  A reported baseline, not the target."""
    import bench_drgbt
    import bench_planning
    rows = []
    try:
        for r in bench_planning.run(64, device=device, cpu=True, max_time=10.0, skip="random obstacles"):
            rows.append({"config": r["scenario"], "planner": r["planner"], "runs": r["runs"], "success": r["success"],
                         "gpu_lockstep_ms_per_run": r["gpu_lockstep"]["per_run_ms"], "gpu_single_run_ms": r["gpu_single"]["per_run_ms"],
                         "cpu_1thread_ms_per_run": r["cpu_1thread"]["per_run_ms"], "cpu_all_threads_ms_per_run": r["cpu_all"]["per_run_ms"],
                         "cpu_cores": r["cpu_all"]["cores"], "cpu_kind": "port (same planner code over the oracle)",
                         "identical_runs_gpu_vs_cpu": r["identical_runs_gpu_vs_cpu"]})
        # the same scenes with 1024 runs per pool: the batch size at which the lockstep pool is meant to be used
        for row, r in zip(rows, bench_planning.run(1024, device=device, cpu=True, max_time=10.0, skip="random obstacles")):
            row["p1024"] = {"gpu_lockstep_ms_per_run": r["gpu_lockstep"]["per_run_ms"], "cpu_all_threads_ms_per_run": r["cpu_all"]["per_run_ms"],
                            "success": r["success"], "identical_runs_gpu_vs_cpu": r["identical_runs_gpu_vs_cpu"]}
        d = bench_drgbt.run(iters=200, obstacles=5, cpu=True, gpu=True, device=device)
        g_arm, c_arm = d["arms"][0], d["arms"][1]
        rows.append({"config": "spatial_22dof DRGBT-shaped replanning loop, 5 moving obstacles (bench_drgbt.py)", "iterations": g_arm["iterations"],
                     "gpu_iteration_ms_median": g_arm["iter_ms"]["median"], "gpu_iteration_ms_p95": g_arm["iter_ms"]["p95"],
                     "cpu_1thread_iteration_ms_median": c_arm["iter_ms"]["median"], "cpu_1thread_iteration_ms_p95": c_arm["iter_ms"]["p95"],
                     "iterations_within_50ms": {"gpu": g_arm["within_50ms"], "cpu": c_arm["within_50ms"]}, "same_run": d.get("same_run")})
    except Exception as e:                                   # the summary must never cost the headline line
        rows.append({"error": repr(e)})
    return rows


# ---- GPU arm -------------------------------------------------------------------------------------------------------------------------
def roofline(eng, rb, prec, precision_name, N, avg_launch_s, pairs_per_launch, work_per_launch, peak32, peak64, clocks, no_cull):
    """Hardware fraction of the dominant kernel.  The kernel is bound by instruction ISSUE (predicated FP32 compares / selects on the ALU
    pipe, integer queue bookkeeping), not by HBM and not by the FMA pipe, so the roofline is issue slots: `achieved` = warp instructions
    executed per launch (hardware counter of the ncu capture of this tree, linear in the number of configurations) / live launch time;
    `peak` = 4 schedulers x 148 SMs x 1 warp instruction per cycle at the SM clock sampled live."""
    peaks, peak_src = measured_peaks()
    L, n = rb.num_links, rb.num_dofs
    esz = 8 if prec == robots.PREC_F64 else 4
    peak_f = peak64 if prec == robots.PREC_F64 else peak32
    c1_flop = N * FLOP_FK + pairs_per_launch * FLOP_PER_PAIR
    alg_bytes = N * (4 * n + esz * (1 + L)) + M_OBSTACLES * 32
    sm_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0
    try:
        import torch
        sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    except Exception:
        sms = 148
    issue_peak = sms * 4 * sm_mhz * 1e6 / 1e9                    # G warp-inst / s
    out = {"bound": "issue", "unit": "Ginst/s", "peak": issue_peak, "peak_source": f"{sms} SMs x 4 schedulers x 1 warp instruction / cycle x {sm_mhz} MHz (SM clock sampled in the timed region)",
           "kernel": ("rvs::distance_kernel<%s,PROFILE>" if no_cull else "rvs::distance_wq_kernel<%s,PROFILE,3> (coherent warps from a counting sort of the batch, distance-field bootstrap, broad-phase scan + warp work queues)") %
                     {"ref32": "double,float", "fast32": "float,float", "f64": "double,double"}[precision_name],
           "avg_launch_ms": 1e3 * avg_launch_s, "configs_per_launch": N,
           "avg_launch_covers": "one rvs_distance_dev call = the three grids of the counting sort (rvs_sort.cuh, ~4 % of the time) + the distance kernel, CUDA events on the engine's stream"}
    prof = None
    if os.path.exists(NCU_PROFILE) and not no_cull and precision_name == "ref32":
        prof = json.load(open(NCU_PROFILE))
    if prof:
        scale = N / prof["configs"]
        inst = prof["warp_inst_executed"] * scale
        flop32 = prof["thread_inst_fp32"]["flop"] * scale
        out.update(achieved=inst / avg_launch_s / 1e9, frac=inst / avg_launch_s / 1e9 / issue_peak,
                   traffic=(prof["dram_bytes_read"] + prof["dram_bytes_write"]) * scale,
                   traffic_note="dram__bytes_read + dram__bytes_write of the distance kernel in the ncu capture quoted below (not measured in this run), scaled to this "
                                "launch; above the algorithmic bytes because the coherent-warp order gathers 24-byte configurations and scatters 4- / 24-byte "
                                "results through 32-byte sectors - under 3 % of the HBM peak either way, the kernel is issue-bound",
                   counters={"source": "profiles/r2/" + os.path.basename(NCU_PROFILE) + " (ncu --set full capture of this kernel on this workload; counts scaled by configurations per launch, "
                                       "time measured live by CUDA events)",
                             "capture": prof["capture"], "capture_configs": prof["configs"], "capture_duration_ms": prof["duration_ms"],
                             "capture_ms_per_1e6_configs": prof["duration_ms"] / prof["configs"] * 1e6, "live_ms_per_1e6_configs": 1e3 * avg_launch_s / N * 1e6,
                             "warp_inst_per_config": prof["warp_inst_executed"] / prof["configs"],
                             "issue_slots_busy_pct_in_capture": prof["issue_slots_busy_pct"], "pipe_alu_pct": prof["pipe_alu_pct"], "pipe_fma_pct": prof["pipe_fma_pct"],
                             "active_threads_per_warp_inst": prof["active_threads_per_warp_inst"], "warps_eligible_per_scheduler": prof["warps_eligible_per_scheduler"],
                             "fp32_arith_tflops": flop32 / avg_launch_s / 1e12, "fp32_arith_frac_of_ffma_peak": flop32 / avg_launch_s / 1e12 / peak32 if peak32 else None,
                             "fp32_arith_note": "FADD + FMUL + 2 FFMA thread instructions (hardware counters): compares, selects and min/max are NOT arithmetic flops, "
                                                "which is why the issue-slot fraction, not this one, is the distance to the hardware limit"})
    else:
        out.update(achieved=None, frac=None, traffic=None,
                   counters={"source": "no ncu capture of this kernel / precision / mode under profiles/r2: issue-slot fraction not available for this run"})
    out["credited"] = {"convention": "C1 (SURVEY.md 8d): 630 flop FK per configuration + 150 flop per pair query the REFERENCE's loop makes", "tflops": c1_flop / avg_launch_s / 1e12,
                       "frac_of_ffma_peak": c1_flop / avg_launch_s / 1e12 / peak_f if peak_f else None, "ffma_peak_tflops": peak_f,
                       "note": "algorithmic credit, NOT hardware utilisation: the broad phase (scan over a distance-sorted table + cut-off) answers ~99 % of the pair "
                               "queries with at most six compares, bit-identically to the un-culled kernel, so this figure exceeds 1"}
    if work_per_launch:
        scan_tests = N * (L * M_OBSTACLES - rb.ground_included)
        out["executed_work"] = {"pair_queries_credited": pairs_per_launch, "scan_survivors": work_per_launch["scan_survivors"], "classified": work_per_launch["classified"],
                                "edges": work_per_launch["edges"], "culled_fraction": 1.0 - work_per_launch["scan_survivors"] / max(1, scan_tests)}
    out["hbm"] = {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / avg_launch_s / 1e9, "peak_gbs": peaks.get("hbm_gbs"), "peak_source": peak_src,
                  "frac": alg_bytes / avg_launch_s / 1e9 / peaks.get("hbm_gbs", 1.0)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--precision", default="ref32", choices=["ref32", "fast32", "f64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the `workloads` points, the isValid figure and the planning-time summary")
    ap.add_argument("--no-cull", action="store_true", help="sequential one-thread-per-configuration kernel (no broad phase)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # libraries (NCCL's version banner, torch warnings) must not pollute stdout: the contract is ONE JSON line there.
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
    esz = 8 if prec == robots.PREC_F64 else 4
    strong = args.scaling == "strong"

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- inputs resident in HBM before the timed region ----
    if strong:
        first, count = shard_range(N_STRONG, rank, world)            # ONE batch, contiguous block per rank (sharding.shard_range)
        nb = 1
        d_q = torch.empty((1, count, n), dtype=torch.float32, device=dev)
        eng.generate_configs_dev(synthetic.SEED_CONFIGS, first, count, d_q[0].data_ptr())
        N = count
    else:
        N = N_CONFIGS
        nb = min(K + W, MAX_RESIDENT_BATCHES)                        # distinct batches; each is 240 MB > L2, so cycling them reuses nothing
        d_q = torch.empty((nb, N, n), dtype=torch.float32, device=dev)
        for s in range(nb):
            eng.generate_configs_dev(synthetic.SEED_CONFIGS, (rank * nb + s) * N, N, d_q[s].data_ptr())
    d_dc = torch.empty(N, dtype=odt, device=dev)
    d_prof = torch.empty((N, L), dtype=odt, device=dev)
    gathered = None
    if strong and world > 1:
        width = max(shard_range(N_STRONG, r, world)[1] for r in range(world))
        d_pack = torch.zeros((width, 1 + L), dtype=odt, device=dev)
        gathered = [torch.empty_like(d_pack) for _ in range(world)] if rank == 0 else None

    def step(s):
        eng.distance_dev(d_q[s % nb].data_ptr(), N, robots.DIST_PROFILE, d_dc.data_ptr(), d_prof.data_ptr())

    def gather_step():
        """strong scaling: (d_c | profile) of every shard to rank 0 over NVLink (NCCL gather = send / recv)"""
        if gathered is None and world == 1:
            return
        d_pack[:N, 0] = d_dc
        d_pack[:N, 1:] = d_prof
        dist.gather(d_pack, gathered, dst=0)

    peak32 = eng.measure_fma_peak(False)
    peak64 = eng.measure_fma_peak(True)
    for s in range(W):
        step(s)
        if strong and world > 1:
            gather_step()
    barrier()
    eng.work_stats(reset=True)
    launches0 = eng.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(K)]
    barrier()
    sampler.mark_begin()
    for s in range(K):
        ev[s][0].record(stream)
        step(W + s)
        ev[s][1].record(stream)
        if strong and world > 1:
            gather_step()
        ev[s][2].record(stream)
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    ms_compute = [a.elapsed_time(b) for a, b, _ in ev]
    ms_gather = [b.elapsed_time(c) for _, b, c in ev]
    t_rank = (sum(ms_compute) + sum(ms_gather)) * 1e-3
    work = eng.work_stats(reset=True)
    pairs_rank = work["pair_queries"]
    launches = eng.launch_count() - launches0
    free_frac = float((d_dc > 0).float().mean().item())

    # ---- parity sample: 2^14 configurations of the LAST timed step against the CPU oracle (the checker, after the timed region) ----
    parity = None
    if rank == 0:
        from oracle import oracle_binding as ob
        ns = min(1 << 14, N)
        last = (W + K - 1) % nb
        q_s = d_q[last, :ns].cpu().numpy()
        d_s, p_s = d_dc[:ns].cpu().numpy(), d_prof[:ns].cpu().numpy()
        orc = ob.Oracle(rb, robots.default_params())
        orc.set_obstacles(mn, mx, g, tol)
        d_o, p_o, _ = orc.distance(q_s, robots.DIST_PROFILE, f64=(prec == robots.PREC_F64), threads=os.cpu_count() or 1)
        parity = {"n": ns, "verdict_mismatch": int(np.count_nonzero((d_s > 0) != (d_o > 0))), "max_abs_d": float(np.abs(d_s - d_o).max()),
                  "max_abs_profile": float(np.abs(p_s - p_o).max()), "bit_identical": bool(np.array_equal(d_s, d_o) and np.array_equal(p_s, p_o)),
                  "checker": "oracle/rvs_oracle.hpp (CPU restatement, pinned bit for bit against the reference's own code compiled here)",
                  "tolerance": "1e-4 m (REF32 / FAST32), 1e-9 m (F64); verdicts identical outside the band"}

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D of the step's configurations, kernel, D2H of d_c + profile ----
    h_q = torch.empty((N, n), dtype=torch.float32).pin_memory()
    h_q.copy_(d_q[W % nb].cpu())
    h_dc = torch.empty(N, dtype=odt).pin_memory()
    h_prof = torch.empty((N, L), dtype=odt).pin_memory()
    e2e_steps = max(3, min(K, 5))
    lib = engine.load_library()
    fn = lib.rvs_distance_f64 if prec == robots.PREC_F64 else lib.rvs_distance
    fptr = ctypes.POINTER(ctypes.c_double if prec == robots.PREC_F64 else ctypes.c_float)

    def e2e_step():
        rc = fn(eng._h, ctypes.cast(h_q.data_ptr(), ctypes.POINTER(ctypes.c_float)), N, robots.DIST_PROFILE,
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
    del h_q, h_dc, h_prof

    # ---- supplementary points (rank 0 of a single-GPU run only; not part of the metric) ----
    verdict, extra, full_mode = None, None, None
    if world == 1 and not args.no_extras and prec != robots.PREC_F64:
        d_v = torch.empty(N, dtype=torch.uint8, device=dev)
        vms = []
        for s in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); eng.is_valid_dev(d_q[s % nb].data_ptr(), N, d_v.data_ptr()); b.record(stream)
            torch.cuda.synchronize()
            vms.append(a.elapsed_time(b))
        vmed = sorted(vms[1:])[1]
        verdict = {"ms_per_launch": vmed, "configs_per_s": N / vmed * 1e3, "collision_pair_queries_per_s": N * synthetic.pair_queries(rb, g) / vmed * 1e3,
                   "valid_fraction": round(float(d_v.float().mean().item()), 4),
                   "kernel": "rvs::valid_wq_kernel (scan + pair / edge phases); nominal pairs L*M per configuration, no early-exit credit"}
        del d_v
        extra = extra_workloads(eng, torch, dev, stream, rb, odt)
        eng.set_obstacles(mn, mx, g, tol)
        full_mode = full_mode_point(torch, dev, stream, rb, prec, local)
    planning_rows = None
    if world == 1 and not args.no_extras and not args.no_cpu_baseline and args.precision == "ref32" and not args.no_cull:
        planning_rows = planning_summary(local)

    # ---- reductions across ranks: time = max, pairs = sum ----
    red = torch.tensor([t_rank, t_e2e, sum(ms_compute) * 1e-3, sum(ms_gather) * 1e-3], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(pairs_rank), float(pairs_e2e), float(launches), float(N)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    t_all, t_e2e_all, t_compute_all, t_gather_all = red.tolist()
    pairs_all, pairs_e2e_all, launches_all, configs_all = tot.tolist()

    if rank == 0:
        avg_launch_s = sum(ms_compute) * 1e-3 / K
        out = {
            "metric": "capsule_obstacle_distance_queries_per_s", "value": pairs_all / t_all, "unit": "pair-queries/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": 1e3 * t_all / K, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": {"ref32": "f32", "fast32": "f32", "f64": "f64"}[args.precision], "data": "synthetic",
            "config": workload_config(world, args.scaling),
            "variant": {"precision": args.precision, "broad_phase": "off" if args.no_cull else "on (results bit-identical to off)"},
            "workload_stats": {"collision_free_fraction": round(free_frac, 4), "pair_queries_per_config": round(pairs_rank / K / N, 1),
                               "configs_per_step_all_ranks": int(configs_all), "timed_region_s": t_all},
            "configs_per_s": configs_all * K / t_all,
            "e2e": {"value": pairs_e2e_all / t_e2e_all, "unit": "pair-queries/s", "h2d_bytes_per_step": N * 4 * n,
                    "d2h_bytes_per_step": N * esz * (1 + L), "steps": e2e_steps, "ms_per_step": 1e3 * t_e2e_all / e2e_steps,
                    "api": "rvs_distance (host pointers, pinned), H2D + kernel + D2H inside the timed region; per rank, all ranks concurrently"},
            "gpu_launches": int(launches_all),
            "parity_sample": parity,
            "roofline": roofline(eng, rb, prec, args.precision, N, avg_launch_s, pairs_rank / K, {k: v / K for k, v in work.items()} if not args.no_cull else None,
                                 peak32, peak64, clocks, args.no_cull),
            "fma_peak_tflops": {"fp32": peak32, "fp64": peak64, "source": "FFMA / DFMA register microbenchmark measured in this run (rvs_measure_fma_peak)"},
            "clocks": clocks,
        }
        if planning_rows is not None:
            out["planning_time_vs_cpu"] = planning_rows
        if strong:
            gb = float(N_STRONG) * esz * (1 + L) * (world - 1) / max(world, 1) / 1e9
            out["strong_scaling"] = {"configs_total": N_STRONG, "compute_ms_per_step": 1e3 * t_compute_all / K, "gather_ms_per_step": 1e3 * t_gather_all / K,
                                     "gathered_gb_per_step": gb, "gather_gbs": (gb / (t_gather_all / K)) if t_gather_all > 0 else None,
                                     "gather": "torch.distributed.gather of (d_c | profile) rows to rank 0 (NCCL send/recv over NVLink), inside the timed region; max over ranks"}
        if verdict:
            out["is_valid"] = verdict
        if extra:
            out["workloads"] = extra
        if full_mode:
            out["full_mode"] = full_mode
        if not args.no_cpu_baseline and world == 1:
            out.update(cpu_baseline())
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def full_mode_point(torch, dev, stream, rb, prec, device):
    """computeDistance with the nearest points of every (link, obstacle) pair (RVS_DIST_FULL, State::nearest_points): N_FULL configurations
    against 51 boxes, device-resident, median of 3 launches after one warm-up; the 24 L M bytes per configuration written per launch
    exceed the L2. Timed twice: the shipped three-launch path (csrc/rvs_distance_full.cuh) and distance_kernel<FULL> (RVS_FULL_SPLIT=0)."""
    from rpmplv2_b200 import engine
    n, L, M = rb.num_dofs, rb.num_links, 51
    mn, mx, g, tol = synthetic.sweep_obstacles(M, "D1", r_min=0.3)
    odt = torch.float32
    d_q = torch.empty((N_FULL, n), dtype=torch.float32, device=dev)
    d_dc = torch.empty(N_FULL, dtype=odt, device=dev)
    d_prof = torch.empty((N_FULL, L), dtype=odt, device=dev)
    d_near = torch.zeros((N_FULL, M, L, 6), dtype=odt, device=dev)
    out = {"configs": N_FULL, "obstacles": M, "result_bytes_per_config": 24 * L * M,
           "note": "nearest points are an output of every pair: nothing is culled; the block of a configuration in collision is unspecified and is not computed by the split path"}
    keep = os.environ.get("RVS_FULL_SPLIT")
    for key, env in (("split", "1"), ("one_kernel", "0")):
        os.environ["RVS_FULL_SPLIT"] = env
        e2 = engine.Engine(rb, robots.default_params(prec), device=device)
        e2.set_stream(stream.cuda_stream)
        e2.set_obstacles(mn, mx, g, tol)
        e2.generate_configs_dev(synthetic.SEED_CONFIGS, 1 << 41, N_FULL, d_q.data_ptr())
        ms = []
        for _ in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0 = e2.launch_count()
            a.record(stream); e2.distance_dev(d_q.data_ptr(), N_FULL, robots.DIST_FULL, d_dc.data_ptr(), d_prof.data_ptr(), d_near.data_ptr()); b.record(stream)
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
            launches = e2.launch_count() - k0
        med = sorted(ms[1:])[1]
        free = float((d_dc > 0).float().mean().item())
        out[key] = {"ms_per_launch": med, "configs_per_s": N_FULL / med * 1e3, "launches": launches,
                    "nearest_points_written_gbs": N_FULL * free * 24 * L * M / med / 1e6}
        out["collision_free_fraction"] = round(free, 4)
        e2.close()
    if keep is None:
        os.environ.pop("RVS_FULL_SPLIT", None)
    else:
        os.environ["RVS_FULL_SPLIT"] = keep
    return out


def extra_workloads(eng, torch, dev, stream, rb, odt):
    """The S1-as-written sweep points and the near-field collision-free point: per launch time, credited pair queries per second,
    what the kernel executed, culled fraction.  N_EXTRA configurations per launch, median of 3 launches after one warm-up; the L2 is
    flushed between launches (these batches are smaller than the L2)."""
    n, L = rb.num_dofs, rb.num_links
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    d_q = torch.empty((N_EXTRA, n), dtype=torch.float32, device=dev)
    eng.generate_configs_dev(synthetic.SEED_CONFIGS, 1 << 40, N_EXTRA, d_q.data_ptr())
    d_dc = torch.empty(N_EXTRA, dtype=odt, device=dev)
    d_prof = torch.empty((N_EXTRA, L), dtype=odt, device=dev)
    out = []
    for name, s in SCENES.items():
        _, mn, mx, g, tol = scene(name)
        eng.set_obstacles(mn, mx, g, tol)
        q, N = d_q, N_EXTRA
        if s["free_only"]:
            eng.distance_dev(d_q.data_ptr(), N_EXTRA, robots.DIST_MIN, d_dc.data_ptr())
            torch.cuda.synchronize()
            q = d_q[d_dc > 0].contiguous()
            N = int(q.shape[0])
        ms = []
        stats = None
        for it in range(4):
            flush.fill_(it)
            eng.work_stats(reset=True)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); eng.distance_dev(q.data_ptr(), N, robots.DIST_PROFILE, d_dc.data_ptr(), d_prof.data_ptr()); b.record(stream)
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
            stats = eng.work_stats(reset=True)
        med = sorted(ms[1:])[1]
        scan_tests = N * synthetic.pair_queries(rb, g)
        out.append({"name": name, "obstacles": s["M"], "sizes": s["dist"], "centres": f"r in [{s['r_min']}, {s['r_max'] or 1.5}] m", "configs": N,
                    "only_collision_free_configs": s["free_only"], "collision_free_fraction": round(float((d_dc[:N] > 0).float().mean().item()), 4),
                    "ms_per_launch": med, "configs_per_s": N / med * 1e3, "credited_pair_queries_per_s": stats["pair_queries"] / med * 1e3,
                    "pair_queries_per_config": stats["pair_queries"] / N, "nominal_pairs_per_config": synthetic.pair_queries(rb, g),
                    "executed_per_config": {"scan_survivors": stats["scan_survivors"] / N, "classified": stats["classified"] / N, "edges": stats["edges"] / N},
                    "culled_fraction": 1.0 - stats["scan_survivors"] / max(1, scan_tests)})
    return out


if __name__ == "__main__":
    main()
