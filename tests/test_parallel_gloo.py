"""world_size-2 gloo tests (CPU) of the multi-rank host logic: shard ranges, result gather, timing reduction."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lstm_ode_bci_b200 import parallel


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 8, 421200, 18944 * 8 + 3):
        for world in (1, 2, 3, 8):
            edges = [parallel.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)
        b, e = parallel.shard_range(n, rank, world)
        got = parallel.gather_shards(full[b:e].clone(), n)
        assert torch.equal(got, full)
        ms = parallel.max_over_ranks(10.0 + rank, "cpu")
        assert ms == 10.0 + world - 1
        # a data-parallel gradient all-reduce equals the single-process gradient of the concatenated batch
        w = torch.ones(4, requires_grad=True)
        x = torch.arange(8, dtype=torch.float32).reshape(2, 4) + 1
        loss = ((x[rank] * w).sum()) / world
        loss.backward()
        g = w.grad.clone()
        dist.all_reduce(g)
        assert torch.allclose(g, x.mean(dim=0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [11, 64])
def test_gather_and_reduce_over_gloo(n):
    port = 29600 + n
    mp.spawn(_worker, args=(2, port, n), nprocs=2, join=True)
