#!/usr/bin/env python
"""profiles/full_small_probe.py — FULL mode around the cross-over between distance_cta_kernel (one CTA per configuration) and the
three-launch split (rvs_distance_full.cuh): device-resident, xArm6, median of 5."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rpmplv2_b200 import engine, robots, synthetic
rb = robots.xarm6()
eng = engine.Engine(rb, robots.default_params(), device=0)
stream = torch.cuda.current_stream(); eng.set_stream(stream.cuda_stream)
for M, dist, rmin in ((51, "D1", 0.3), (20, "D2", 0.45), (300, "D2", 0.9)):
    mn, mx, g, tol = synthetic.sweep_obstacles(M, dist, r_min=rmin)
    eng.set_obstacles(mn, mx, g, tol)
    for N in (512, 1024, 2048, 4096, 8192, 16384, 32768):
        d_q = torch.empty((N, 6), dtype=torch.float32, device="cuda")
        eng.generate_configs_dev(synthetic.SEED_CONFIGS, 0, N, d_q.data_ptr())
        d_dc = torch.empty(N, dtype=torch.float32, device="cuda"); d_p = torch.empty((N, 6), dtype=torch.float32, device="cuda")
        d_n = torch.empty((N, M, 6, 6), dtype=torch.float32, device="cuda")
        ms = []
        for _ in range(7):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0 = eng.launch_count()
            a.record(stream); eng.distance_dev(d_q.data_ptr(), N, robots.DIST_FULL, d_dc.data_ptr(), d_p.data_ptr(), d_n.data_ptr()); b.record(stream); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
            k = eng.launch_count() - k0
        print("M", M, "N", N, "FULL ms %.4f" % float(np.median(ms[2:])), "launches", k, flush=True)
