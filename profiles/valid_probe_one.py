#!/usr/bin/env python
"""profiles/valid_probe_one.py <dist> <r_min> <M> — a few isValid launches on one S1 scene (for an ncu capture of valid_wq_kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rpmplv2_b200 import engine, robots, synthetic
dist, rmin, M = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
rb = robots.xarm6()
eng = engine.Engine(rb, robots.default_params(), device=0)
stream = torch.cuda.current_stream(); eng.set_stream(stream.cuda_stream)
N = 1 << 20
d_q = torch.empty((N, 6), dtype=torch.float32, device="cuda"); d_v = torch.empty(N, dtype=torch.uint8, device="cuda")
eng.generate_configs_dev(synthetic.SEED_CONFIGS, 0, N, d_q.data_ptr())
eng.set_obstacles(*synthetic.sweep_obstacles(M, dist, r_min=rmin))
for _ in range(4):
    eng.is_valid_dev(d_q.data_ptr(), N, d_v.data_ptr())
torch.cuda.synchronize()
print("valid", d_v.float().mean().item())
