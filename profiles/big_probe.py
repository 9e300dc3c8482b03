#!/usr/bin/env python
"""profiles/big_probe.py — the large-N rows of SURVEY.md 8(d) S1: 10^7 and 10^8 xArm6 configurations x 1000 boxes (PROFILE / VERDICT),
device-resident inputs from the stateless generator, one launch each, CUDA events; properties checked instead of an oracle run
(d_c == min(profile) where collision-free; d_c == 0 <=> profile has a non-positive entry; verdict false => d_c == 0 is NOT implied, Q-list)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rpmplv2_b200 import engine, robots, synthetic
rb = robots.xarm6()
eng = engine.Engine(rb, robots.default_params(), device=0)
stream = torch.cuda.current_stream(); eng.set_stream(stream.cuda_stream)
mn, mx, g, tol = synthetic.sweep_obstacles(1000, "D2", r_min=0.9)
eng.set_obstacles(mn, mx, g, tol)
for N in (10_000_000, 100_000_000):
    d_q = torch.empty((N, 6), dtype=torch.float32, device="cuda")
    eng.generate_configs_dev(synthetic.SEED_CONFIGS, 0, N, d_q.data_ptr())
    d_dc = torch.empty(N, dtype=torch.float32, device="cuda"); d_p = torch.empty((N, 6), dtype=torch.float32, device="cuda")
    d_v = torch.empty(N, dtype=torch.uint8, device="cuda")
    eng.pair_count(reset=True)
    ms = []
    for _ in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); eng.distance_dev(d_q.data_ptr(), N, robots.DIST_PROFILE, d_dc.data_ptr(), d_p.data_ptr()); b.record(stream); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    pairs = eng.pair_count(reset=True) / 2
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); eng.is_valid_dev(d_q.data_ptr(), N, d_v.data_ptr()); b.record(stream); torch.cuda.synchronize(); msv = a.elapsed_time(b)
    free = d_dc > 0
    ok_min = bool((d_p[free].min(dim=1).values == d_dc[free]).all().item())
    ok_hit = bool(((d_p[~free] <= 0).any(dim=1)).all().item())
    print(f"N {N:.0e} PROFILE ms {ms[1]:.1f} pair-queries/s {pairs / ms[1] * 1e3:.3e} configs/s {N / ms[1] * 1e3:.3e} free {free.float().mean().item():.4f} "
          f"min(profile)==d_c {ok_min} collision=>profile<=0 {ok_hit} | VERDICT ms {msv:.1f} configs/s {N / msv * 1e3:.3e} valid {d_v.float().mean().item():.4f}", flush=True)
    del d_q, d_dc, d_p, d_v
    torch.cuda.empty_cache()
