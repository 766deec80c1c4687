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


def test_host_side_helpers_cpu():
    """Host logic that needs no GPU: the multi-threaded staging copy and the RK4 stage-time grid of solve_with_modulation."""
    import numpy as np
    import torch
    from lstm_ode_bci_b200 import integration, ode
    src = torch.from_numpy(np.random.default_rng(1).standard_normal((4099, 64, 17), dtype=np.float32))   # > 4 Mi elements: threaded path
    dst = torch.empty_like(src)
    integration._staging_copy(dst, src)
    assert torch.equal(dst, src)
    small = torch.arange(12, dtype=torch.float32).reshape(3, 4)
    out = torch.empty_like(small)
    integration._staging_copy(out, small)
    assert torch.equal(out, small)
    tn = ode.modulation_nodes(5.0, 25.0, 20, 4)
    assert len(tn) == 2 * 4 * 19 + 1 and tn[0] == 5.0 and tn[-1] == 25.0
    h = (25.0 - 5.0) / 19 / 4
    assert np.allclose(np.diff(tn), h / 2)
    assert np.allclose(tn[::8], np.linspace(5.0, 25.0, 20))       # every 2*substeps-th node is an output time
