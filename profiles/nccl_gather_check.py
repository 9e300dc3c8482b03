#!/usr/bin/env python
"""profiles/nccl_gather_check.py — the sharded sweep with a REAL NCCL gather (one rank per GPU, torchrun): every rank computes its
contiguous block on its own GPU from the stateless generator, the results are gathered to rank 0 over NCCL (sharding.gather_rows on
device tensors) and compared bit for bit with rank 0 computing the whole batch alone.  (The gloo twin runs on CPU in tests/.)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from rpmplv2_b200 import engine, robots, sharding, synthetic

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
rb = robots.xarm6()
eng = engine.Engine(rb, robots.default_params(), device=local)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
eng.set_obstacles(*synthetic.sweep_obstacles(1000, "D2", r_min=0.9))
N = 3_000_001


def compute(first, count):
    q = torch.empty((count, 6), dtype=torch.float32, device=dev)
    d = torch.empty(count, dtype=torch.float32, device=dev); p = torch.empty((count, 6), dtype=torch.float32, device=dev)
    v = torch.empty(count, dtype=torch.uint8, device=dev)
    eng.generate_configs_dev(synthetic.SEED_CONFIGS, first, count, q.data_ptr())
    eng.distance_dev(q.data_ptr(), count, robots.DIST_PROFILE, d.data_ptr(), p.data_ptr())
    eng.is_valid_dev(q.data_ptr(), count, v.data_ptr())
    torch.cuda.synchronize()
    return d, p, v


torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
res = sharding.run_sharded(compute, N, rank, world, dst=0)
torch.cuda.synchronize(); dist.barrier()
t1 = time.perf_counter()
if rank == 0:
    full = compute(0, N)
    ok = all(bool((a == b).all().item()) for a, b in zip(res, full))
    print(f"world {world}: sharded sweep + NCCL gather of {N} configurations x 1000 boxes in {1e3 * (t1 - t0):.1f} ms; "
          f"gathered == single-GPU result bit for bit: {ok}; free {float((full[0] > 0).float().mean()):.4f}")
    assert ok
dist.barrier()
dist.destroy_process_group()
