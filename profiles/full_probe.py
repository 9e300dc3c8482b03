#!/usr/bin/env python
"""profiles/full_probe.py — FULL-mode (nearest points) throughput, device-resident, xArm6 x 51 boxes, N = 131072."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rpmplv2_b200 import engine, robots, synthetic
rb = robots.xarm6()
eng = engine.Engine(rb, robots.default_params(), device=0)
stream = torch.cuda.current_stream(); eng.set_stream(stream.cuda_stream)
for N, M, dist, rmin in ((131072, 51, "D1", 0.3), (1048576, 51, "D1", 0.3), (16384, 51, "D1", 0.3), (2048, 51, "D1", 0.3), (131072, 6, "D1", 0.3), (131072, 51, "D2", 0.9), (32768, 300, "D2", 0.9)):
    mn, mx, g, tol = synthetic.sweep_obstacles(M, dist, r_min=rmin)
    eng.set_obstacles(mn, mx, g, tol)
    d_q = torch.empty((N, 6), dtype=torch.float32, device="cuda")
    eng.generate_configs_dev(synthetic.SEED_CONFIGS, 0, N, d_q.data_ptr())
    d_dc = torch.empty(N, dtype=torch.float32, device="cuda"); d_p = torch.empty((N, 6), dtype=torch.float32, device="cuda")
    d_n = torch.empty((N, M, 6, 6), dtype=torch.float32, device="cuda")
    for mode, name, srt in ((robots.DIST_FULL, "FULL", 1), (robots.DIST_FULL, "FULL unsorted", 0), (robots.DIST_PROFILE, "PROFILE", 1)):
        eng.set_config_sort(bool(srt))
        ms = []
        for _ in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); eng.distance_dev(d_q.data_ptr(), N, mode, d_dc.data_ptr(), d_p.data_ptr(), d_n.data_ptr()); b.record(stream); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
        t = float(np.median(ms[1:]))
        if name == "FULL" and N == 131072: print("   collision-free fraction %.3f" % float((d_dc > 0).float().mean()), "RVS_FULL_SPLIT", os.environ.get("RVS_FULL_SPLIT", "1"), flush=True)
        print("N", N, "M", M, dist, rmin, name, "ms %.3f" % t, "configs/s %.3g" % (N / t * 1e3), "GB/s written %.0f" % (N * M * 6 * 24 / t / 1e6) if mode == robots.DIST_FULL else "", flush=True)
