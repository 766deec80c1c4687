"""Multi-GPU plumbing (SURVEY.md §8 e): windows and trajectories are independent, so ranks own contiguous
shards and never exchange data on the hot path; the only collectives are the gradient all-reduce of the
training step (train.FusedTrainer) and the final result gather below.  Works on any torch.distributed backend
(nccl on the GPUs; the CPU tests run it over gloo)."""
import ctypes as C

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


def _ode_and_forecast(integration, probs, n_total, horizons, ode_params, group, want_traj=True):
    """Stages 2-4 of config 5 on this rank's probabilities: coupled ODE (06 path), gather of the (N,2) probabilities, 08-style
    forecast of this rank's share of the N - max(h) forecast origins."""
    from . import ops
    from .integration import _forecast_device
    from .ode import solve_ensemble
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n = probs.shape[0]
    traj, final, _ = solve_ensemble(n, p_open=probs[:, 0].contiguous(), p_closed=probs[:, 1].contiguous(),
                                    base_rates=integration.base_params, alpha=integration.coupling_strength, y0_mode="probs06",
                                    coupling=True, style="ref06", mode="rk4", t_end=20.0, n_points=20,
                                    substeps=integration.substeps, want_traj=want_traj, device=probs.device)
    pred, cls = ops.ode_classify(final, True, True)
    all_probs = gather_shards(probs, n_total, group)
    m = n_total - max(horizons)
    b, e = shard_range(max(m, 0), rank, world)
    fc = None
    if e > b:
        prm = ode_params if ode_params is not None else integration.base_params
        fc = _forecast_device(all_probs[b:e, 1].contiguous(), prm, max(horizons), list(horizons), probs.device, integration.substeps)
    return {"traj": traj, "final": final, "pred": pred, "cls": cls, "probs": all_probs, "forecast": fc, "forecast_range": (b, e)}


def forecast_pipeline_sharded(integration, X_local, n_total, horizons=(5, 10, 20), ode_params=None, group=None):
    """Config 5: each rank classifies its contiguous window shard and integrates its coupled trajectories;
    the 08-style forecast needs probs[i+h] across shard edges, so the (N,2) probabilities are gathered first
    (3.4 MB for 421 200 windows) and the forecast ODE stage is sharded again.  Returns rank-local tensors plus the
    gathered probabilities."""
    from .integration import _lstm_probs_device
    probs, _ = _lstm_probs_device(integration.lstm_model, X_local, None, False, X_local.device, autocast=True)
    return _ode_and_forecast(integration, probs, n_total, tuple(horizons), ode_params, group)


def forecast_pipeline_from_recordings(integration, host_recording_batches, n_total, horizons=(5, 10, 20), ode_params=None,
                                      group=None, raw=True, want_traj=True, **stream_kw):
    """Config 5 end to end FROM THE HOST: this rank's recordings are streamed to its GPU (copy stream, double-buffered) and the
    windows are cut on the device -- raw=True: raw (R, C, n) recordings -> band-pass + z-score + windowing (`bci_preprocess`,
    02_preprocessing.py:114-180) -> BiLSTM; raw=False: normalised sample-major (R, S, C) recordings (fp32 or bf16) read in
    place by the input projection -- then coupling + ODE, the probability gather and the 08-style forecast as in
    `forecast_pipeline_sharded`.  `n_total` = windows over all ranks (ranks own contiguous recording ranges, so window order is
    rank order)."""
    import torch
    from .integration import stream_raw_recordings, stream_recordings
    model = integration.lstm_model
    with torch.autocast("cuda", dtype=torch.bfloat16):           # where the reference's predict_batch autocasts (06:348-351)
        if raw:
            parts = [p for p, _ in stream_raw_recordings(model, host_recording_batches, device=integration.device, **stream_kw)]
        else:
            parts = [p for p, _ in stream_recordings(model, host_recording_batches, device=integration.device, **stream_kw)]
    probs = parts[0] if len(parts) == 1 else torch.cat(parts)
    return _ode_and_forecast(integration, probs, n_total, tuple(horizons), ode_params, group, want_traj=want_traj)


class _DevArray:
    """Zero-copy torch view of a raw device allocation owned by the C library (via __cuda_array_interface__)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 3}


class P2PComm:
    """Peer-memory communicator of the fused optimizer step (`bci_fused_step`, csrc/comm_p2p.cu).

    Each rank's gradient bucket is allocated by the library and exported as a CUDA IPC handle; the handles travel
    through torch.distributed (any backend) once at construction.  After that a training step involves no NCCL
    call: the reduction happens inside the optimizer kernels over NVLink peer loads.  GPU only (no CPU path)."""

    def __init__(self, n_floats, group=None, device=None):
        from . import _native as N
        self._N = N
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n = int(n_floats)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.ptr = C.c_void_p(0)
        with torch.cuda.device(self.device):
            N.check(N.lib().bci_comm_create(self.rank, self.world, self.n, C.byref(self.ptr)))
            if self.world > 1:
                blob = C.create_string_buffer(N.COMM_HANDLE_BYTES)
                N.check(N.lib().bci_comm_export(self.ptr, blob))
                blobs = [None] * self.world
                dist.all_gather_object(blobs, bytes(blob.raw), group=group)
                N.check(N.lib().bci_comm_connect(self.ptr, C.create_string_buffer(b"".join(blobs), N.COMM_HANDLE_BYTES * self.world)))
            bp, bn = C.c_void_p(0), C.c_int64(0)
            N.check(N.lib().bci_comm_bucket(self.ptr, C.byref(bp), C.byref(bn)))
        self.bucket = torch.as_tensor(_DevArray(bp.value, bn.value), device=self.device)
        if self.world > 1:
            dist.barrier(group=group)      # every rank has opened every handle before anyone steps

    def fused_step(self, p, m, v, lr, betas, eps, weight_decay, step, max_norm, norm_out=None):
        N = self._N
        from .ops import _ptr, _stream
        N.check(N.lib().bci_fused_step(self.ptr, _ptr(p), _ptr(m), _ptr(v), lr, betas[0], betas[1], eps, weight_decay,
                                       int(step), max_norm, _ptr(norm_out), _stream()))

    def close(self):
        """Frees the bucket: every tensor view of it (`.bucket`, a trainer's gradient views) is invalid afterwards."""
        if self.ptr:
            torch.cuda.synchronize(self.device)
            self.bucket = None
            self._N.lib().bci_comm_destroy(self.ptr)
            self.ptr = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
