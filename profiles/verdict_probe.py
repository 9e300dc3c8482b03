#!/usr/bin/env python
"""profiles/verdict_probe.py — isValid (VERDICT) timing on the S1 scenes (device-resident configurations, CUDA events)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rpmplv2_b200 import engine, robots, synthetic
rb = robots.xarm6()
eng = engine.Engine(rb, robots.default_params(), device=0)
stream = torch.cuda.current_stream(); eng.set_stream(stream.cuda_stream)
N = 1 << 20
d_q = torch.empty((N, 6), dtype=torch.float32, device="cuda"); d_v = torch.empty(N, dtype=torch.uint8, device="cuda")
eng.generate_configs_dev(synthetic.SEED_CONFIGS, 0, N, d_q.data_ptr())
for dist, rmin in (("D2", 0.9), ("D1", 0.0), ("D2", 0.0)):
    for M in (10, 100, 1000):
        eng.set_obstacles(*synthetic.sweep_obstacles(M, dist, r_min=rmin))
        for cull in (1, 0):
            eng.set_culling(bool(cull))
            ms = []
            for _ in range(4):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream); eng.is_valid_dev(d_q.data_ptr(), N, d_v.data_ptr()); b.record(stream); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
            print(dist, rmin, M, "wq" if cull else "seq", "ms %.3f" % np.median(ms[1:]), "valid %.4f" % d_v.float().mean().item(), flush=True)
