"""Multi-GPU host logic on CPU: world_size-2 gloo processes shard a configuration batch, each computes its block (the oracle stands
in for the per-GPU engine here — this is a test of the partition + gather, not of the kernels) and rank 0 gathers d_c / profile."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from rpmplv2_b200 import robots, sharding, synthetic


def test_shard_ranges_partition_the_batch():
    for N in (0, 1, 7, 1000, 2 ** 20 + 3):
        for world in (1, 2, 3, 8):
            blocks = [sharding.shard_range(N, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == N
            assert all(blocks[r][0] + blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def _worker(rank, world, port, N, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle_binding as ob
        import scenes
        rb, mn, mx, g, tol = scenes.xarm6_scene(20, "D2", 0.4)
        o = ob.Oracle(rb, robots.default_params())
        o.set_obstacles(mn, mx, g, tol)

        def compute(first, count):                       # stateless inputs: no scatter needed
            q = synthetic.configs(rb, count, first=first)
            d, p, _ = o.distance(q, robots.DIST_PROFILE)
            return d, p, o.is_valid(q)

        res = sharding.run_sharded(compute, N, rank, world, dst=0)
        local = sharding.run_sharded(compute, N, rank, world, dst=None)
        assert local[0].shape[0] == sharding.shard_range(N, rank, world)[1]
        if rank == 0:
            np.savez(out_path, d=res[0], p=res[1], v=res[2])
        else:
            assert res is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N", [1001, 64])
def test_two_rank_gloo_sharded_sweep_equals_single_process(ob, tmp_path, N):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, port, N, out), nprocs=2, join=True)
    import scenes
    rb, mn, mx, g, tol = scenes.xarm6_scene(20, "D2", 0.4)
    o = ob.Oracle(rb, robots.default_params())
    o.set_obstacles(mn, mx, g, tol)
    q = synthetic.configs(rb, N)
    d, p, _ = o.distance(q, robots.DIST_PROFILE)
    got = np.load(out)
    assert got["d"].shape == (N,) and (got["d"] == d).all() and (got["p"] == p).all() and (got["v"] == o.is_valid(q)).all()
