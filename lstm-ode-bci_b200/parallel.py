"""Multi-GPU plumbing (SURVEY.md §8 e): windows and trajectories are independent, so ranks own contiguous
shards and never exchange data on the hot path; the only collectives are the gradient all-reduce of the
training step (train.FusedTrainer) and the final result gather below.  Works on any torch.distributed backend
(nccl on the GPUs; the CPU tests run it over gloo)."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous [begin, end) of n items for `rank`; sizes differ by at most 1, earlier ranks get the extra."""
    base, extra = divmod(int(n), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_shards(local, n_total, group=None):
    """All-gather variable-length row shards (as produced by shard_range) into the full (n_total, ...) tensor."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    longest = max(e - b for b, e in sizes)
    pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([parts[r][: sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0)


def max_over_ranks(value, device, group=None):
    """Device-timed milliseconds -> max over ranks (the number bench.py reports)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def forecast_pipeline_sharded(integration, X_local, n_total, horizons=(5, 10, 20), ode_params=None, group=None):
    """Config 5: each rank classifies its contiguous window shard and integrates its coupled trajectories;
    the 08-style forecast needs probs[i+h] across shard edges, so the (N,2) probabilities are gathered first
    (3.4 MB for 421 200 windows) and the forecast ODE stage is sharded again.  Returns rank-local tensors plus the
    gathered probabilities."""
    from . import ops
    from .integration import _forecast_device
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    traj, probs, final, pred, cls = integration.predict_batch_device(X_local)
    all_probs = gather_shards(probs, n_total, group)
    m = n_total - max(horizons)
    b, e = shard_range(max(m, 0), rank, world)
    fc = None
    if e > b:
        prm = ode_params if ode_params is not None else integration.base_params
        fc = _forecast_device(all_probs[b:e, 1].contiguous(), prm, max(horizons), list(horizons), X_local.device, integration.substeps)
    return {"traj": traj, "final": final, "pred": pred, "cls": cls, "probs": all_probs, "forecast": fc, "forecast_range": (b, e)}
