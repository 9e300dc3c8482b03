import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import torch
from rpmplv2_b200 import engine, robots, synthetic
rb = robots.xarm6()
eng = engine.Engine(rb, robots.default_params(), device=0)
stream = torch.cuda.current_stream(); eng.set_stream(stream.cuda_stream)
N, M = int(os.environ.get("RVS_FULL_NCU_CONFIGS", 131072)), 51
mn, mx, g, tol = synthetic.sweep_obstacles(M, "D1", r_min=0.3)
eng.set_obstacles(mn, mx, g, tol)
d_q = torch.empty((N, 6), dtype=torch.float32, device="cuda")
eng.generate_configs_dev(synthetic.SEED_CONFIGS, 0, N, d_q.data_ptr())
d_dc = torch.empty(N, dtype=torch.float32, device="cuda"); d_p = torch.empty((N, 6), dtype=torch.float32, device="cuda")
d_n = torch.empty((N, M, 6, 6), dtype=torch.float32, device="cuda")
for _ in range(3):
    eng.distance_dev(d_q.data_ptr(), N, robots.DIST_FULL, d_dc.data_ptr(), d_p.data_ptr(), d_n.data_ptr()); torch.cuda.synchronize()
