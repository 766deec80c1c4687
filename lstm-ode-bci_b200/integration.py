"""Host-side mirrors of the reference's LSTM -> coupling -> ODE callers.

  LSTMODEIntegration           06_lstm_ode_integration.py:183-406
  get_three_state_probabilities 10_three_state_probabilities.py:204-290
  prob_to_ode_state / predict_trajectory / multistep_forecast / rolling_forecast_evaluation
                               08_forecasting.py:149-153,215-289,346-392

Same names, arguments, return shapes/dtypes; but the per-sample host loops of the reference
(06:372-401, 10:245-273, 08:264-282) become ONE ODE-ensemble launch, and the LSTM
probabilities never leave the device between the two stages.
"""
import os

import numpy as np
import torch

from . import _native as N
from . import ops
from .ode import CognitiveStateODE, solve_ensemble, _dev
from .synth import RATE_ORDER


def _staging_threads():
    """BCI_STAGING_THREADS, else this rank's share of the host cores (a core streams ~7 GB/s whatever the instruction mix: the
    copy scales with threads until the memory controllers saturate), at most 32."""
    env = os.environ.get("BCI_STAGING_THREADS")
    if env:
        return max(int(env), 1)
    try:
        cores = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        cores = os.cpu_count() or 1
    local_world = max(int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1), 1)
    return max(1, min(32, cores // local_world))


def _staging_copy(dst, src):
    """pageable -> pinned staging copy on several host threads (`bci_host_stage`: streaming stores, optional fp32 -> bf16 narrowing;
    ctypes releases the GIL).  One thread moves ~7 GB/s, the PCIe link takes ~53 GB/s from the pinned buffer: this leg bounds the
    drop-in callers, which is why it is native."""
    workers = _staging_threads()
    if not (src.is_contiguous() and dst.is_contiguous() and src.dtype == torch.float32 and dst.dtype in (torch.float32, torch.bfloat16)):
        dst.copy_(src)          # strided or non-fp32 host input: torch's own copy (still host -> pinned staging, same data path)
        return
    N.check(N.lib().bci_host_stage(dst.data_ptr(), src.data_ptr(), src.numel(), 1 if dst.dtype == torch.bfloat16 else 0, max(workers, 1)))


def stream_lstm_probs(lstm_model, host_batches, device=None, chunk=None, want_attn=False):
    """Pipelined inference over a stream of HOST batches (the reference copies each batch synchronously, 06:346).

    host_batches: iterable of CPU float32 tensors / numpy arrays (n_i, T, C).  Pinned tensors are DMA-ed in
    place; pageable ones go through pinned staging buffers.  Yields, in order, (probs, attention-or-None) as
    CUDA tensors enqueued on the current stream (use them on that stream, or synchronise).  The H2D copy of batch i+1 (copy stream, `chunk`-window
    pieces) overlaps the kernels of batch i (current stream); `chunk` defaults to one full wave of recurrence CTAs."""
    dev = _dev(device)
    lstm_model.eval()
    if chunk is None:
        chunk = ops.lstm_chunk_windows(lstm_model._engine(lstm_model._precision_now()))
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    slots = [dict(buf=None, free=None, pin=None) for _ in range(2)]
    # bf16 engine + pageable input: the staging copy narrows to bf16 (the rounding the input projection applies on load anyway, so
    # no result bit changes) and the link carries half the bytes; the windows are then read as one "recording" with step = seq_len
    narrow_ok = lstm_model._precision_now() == "bf16" and os.environ.get("BCI_STAGING_BF16", "1") != "0"

    def submit(k, xb):
        sl = slots[k & 1]
        xh = torch.as_tensor(xb)
        if xh.dtype != torch.float32:
            xh = xh.float()
        n = xh.shape[0]
        direct = xh.is_pinned() and xh.is_contiguous()
        dt = torch.bfloat16 if (narrow_ok and not direct and xh.dim() == 3) else torch.float32
        if sl["buf"] is None or sl["buf"].shape[0] < n or sl["buf"].shape[1:] != xh.shape[1:] or sl["buf"].dtype != dt:
            sl["buf"] = torch.empty(tuple(xh.shape), dtype=dt, device=dev)
            # the block may have just been freed by main-stream work that is still in flight (the caching allocator hands it
            # out in main-stream order): the copy stream must not write into it before that work has finished
            copy_stream.wait_stream(main)
        if sl["free"] is not None:
            copy_stream.wait_event(sl["free"])           # kernels of the batch that last used this buffer are done
        if not direct:
            if sl["pin"] is None or sl["pin"].shape[0] < n or sl["pin"].shape[1:] != xh.shape[1:] or sl["pin"].dtype != dt:
                sl["pin"] = torch.empty(tuple(xh.shape), dtype=dt, pin_memory=True)   # (.pin_memory() would copy a pageable tensor first)
            if sl["free"] is not None:
                sl["free"].synchronize()
        events = []
        # pageable input: staged and copied in eighths of a pass, so the DMA of one eighth overlaps the staging of the next (the
        # compute still runs per pass, after the pass's last eighth has landed)
        sub = chunk if direct else max((chunk + 7) // 8, 1)
        for i in range(0, n, chunk):
            m = min(chunk, n - i)
            ev = None
            for j in range(i, i + m, sub):
                mj = min(sub, i + m - j)
                src = xh[j:j + mj]
                if not direct:
                    _staging_copy(sl["pin"][j:j + mj], src)
                    src = sl["pin"][j:j + mj]
                with torch.cuda.stream(copy_stream):
                    sl["buf"][j:j + mj].copy_(src, non_blocking=True)
            with torch.cuda.stream(copy_stream):
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            events.append((i, m, ev))
        return sl, n, events

    def compute(sl, n, events):
        probs = torch.empty((n, lstm_model.num_classes), device=dev, dtype=torch.float32)
        attn = None
        with torch.no_grad():
            for i, m, ev in events:
                main.wait_event(ev)
                xb = sl["buf"][i:i + m]
                if xb.dtype == torch.bfloat16:
                    T = xb.shape[1]
                    p = lstm_model.predict_proba_recordings(xb.view(1, m * T, xb.shape[2]), T, T, return_attention=want_attn)
                else:
                    p = lstm_model.predict_proba(xb, return_attention=want_attn)
                if want_attn:
                    p, a = p
                    if attn is None:
                        attn = torch.empty((n, a.shape[1]), device=dev, dtype=torch.float32)
                    attn[i:i + m] = a
                probs[i:i + m] = p
        sl["free"] = torch.cuda.Event()
        sl["free"].record(main)
        return probs, attn

    it = iter(host_batches)
    k = 0
    try:
        pending = submit(k, next(it))
    except StopIteration:
        return
    while pending is not None:
        out = compute(*pending)                           # batch k's kernels are queued (they wait on its copy events) ...
        try:
            nxt = submit(k + 1, next(it))                 # ... and run while the host stages batch k + 1 and queues its copies
        except StopIteration:
            nxt = None
        yield out
        pending = nxt
        k += 1


class _H2DRing:
    """Two device buffers fed by a copy stream: the H2D copy of batch k+1 overlaps the kernels of batch k (main stream).
    Pinned contiguous host tensors are DMA-ed in place; pageable ones go through a pinned staging buffer per slot."""

    def __init__(self, dev):
        self.dev = dev
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.main = torch.cuda.current_stream(dev)
        self.slots = [dict(buf=None, free=None, pin=None) for _ in range(2)]

    def submit(self, k, host):
        sl = self.slots[k & 1]
        if sl["buf"] is None or sl["buf"].shape != host.shape or sl["buf"].dtype != host.dtype:
            sl["buf"] = torch.empty(tuple(host.shape), dtype=host.dtype, device=self.dev)
            # a block the caching allocator just recycled may still be in use by main-stream work in flight
            self.copy_stream.wait_stream(self.main)
        if sl["free"] is not None:
            self.copy_stream.wait_event(sl["free"])          # kernels of the batch that last used this buffer are done
        src = host
        if not (host.is_pinned() and host.is_contiguous()):
            if sl["pin"] is None or sl["pin"].shape != host.shape or sl["pin"].dtype != host.dtype:
                sl["pin"] = torch.empty(tuple(host.shape), dtype=host.dtype, pin_memory=True)
            if sl["free"] is not None:
                sl["free"].synchronize()
            _staging_copy(sl["pin"], host)
            src = sl["pin"]
        with torch.cuda.stream(self.copy_stream):
            sl["buf"].copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return sl, ev

    def acquire(self, token):
        sl, ev = token
        self.main.wait_event(ev)
        return sl["buf"]

    def release(self, token):
        sl, _ = token
        sl["free"] = torch.cuda.Event()
        sl["free"].record(self.main)

    def run(self, host_batches, prepare, compute):
        it = iter(host_batches)
        k = 0
        try:
            pending = self.submit(k, prepare(next(it)))
        except StopIteration:
            return
        while pending is not None:
            out = compute(self.acquire(pending))              # batch k's kernels are queued ...
            self.release(pending)
            try:
                nxt = self.submit(k + 1, prepare(next(it)))   # ... and run while the host stages batch k + 1 / its copy is queued
            except StopIteration:
                nxt = None
            yield out
            pending = nxt
            k += 1


def stream_recordings(lstm_model, host_batches, seq_len=256, step=128, device=None, want_attn=False):
    """Pipelined inference straight from normalised RECORDINGS on the host: the windows are never materialised.

    The reference writes every window to disk and host memory (create_sequences, 02_preprocessing.py:157-180: 50 % overlap, so
    each sample is stored twice) and copies fp32 windows to the GPU (06:346: 62 464 B per window).  Here the host keeps what 02
    has just before that step -- the band-passed, z-scored recording, sample-major (S, C) -- and the input projection cuts the
    windows out of it on the device (`bci_lstm_forward_view`): 31 232 B per window in float32, 15 616 B in bfloat16 (the bf16
    tensor-core mode rounds x to bf16 on load anyway: same result bits, a quarter of the reference's bytes).

    host_batches: iterable of CPU tensors / numpy arrays (R_i, S, C), float32 or torch.bfloat16.  Yields, in order,
    (probs (R_i * n_seq, 2), attention-or-None) as CUDA tensors enqueued on the current stream; n_seq = (S - seq_len) // step + 1,
    windows ordered recording by recording as create_sequences emits them."""
    dev = _dev(device)
    lstm_model.eval()

    def prepare(rb):
        rh = torch.as_tensor(rb)
        if rh.dtype not in (torch.float32, torch.bfloat16):
            rh = rh.float()
        if rh.dim() != 3:
            raise N.BciError(-1, "a recording batch must be (R, S, C), got %s" % (tuple(rh.shape),))
        return rh

    def compute(buf):
        out = lstm_model.predict_proba_recordings(buf, seq_len=seq_len, step=step, return_attention=want_attn)
        return out if want_attn else (out, None)

    yield from _H2DRing(dev).run(host_batches, prepare, compute)


def stream_raw_recordings(lstm_model, host_batches, lowcut=1.0, highcut=45.0, fs=500, order=4, seq_len=256, overlap=0.5,
                          normalization_params=None, device=None):
    """Pipelined inference from RAW recordings on the host (what mne's raw.get_data() returns, 02_preprocessing.py:200):
    H2D of (R_i, C, n) float32/float64 batches on a copy stream -> band-pass filtfilt + per-channel z-score + 50 %-overlap
    windowing on the device (`bci_preprocess`, 02:114-180) -> BiLSTM forward.  31 232 B per window cross PCIe in float32
    instead of the 62 464 B of materialised fp32 windows, and the filter runs at GPU rate.  Yields (probs (R_i * n_seq, 2),
    dict(mean, std)) per batch, CUDA tensors on the current stream."""
    from . import preprocessing as pp
    dev = _dev(device)
    lstm_model.eval()
    b, a, zi, padlen = pp.design_bandpass(lowcut, highcut, fs, order)
    mean = std = None
    if normalization_params:
        mean, std = normalization_params["mean"], normalization_params["std"]

    def prepare(rb):
        rh = torch.as_tensor(rb)
        if rh.dtype not in (torch.float32, torch.float64):
            rh = rh.double()
        if rh.dim() == 2:
            rh = rh[None]
        return rh

    # (Tried: preprocessing of batch k+1 on its own stream beside the BiLSTM of batch k -- the filter is a latency-bound fp64 recursion
    # with few threads.  Its blocks delay the 4-CTA cluster launches of the LSTM: config 5 ran at 0.76 M instead of 1.05 M windows/s.
    # The stages of one batch therefore run back to back; only the H2D copy of the next batch overlaps them.)
    def compute(buf):
        out = pp.preprocess_recordings(buf, b, a, zi, padlen, seq_len, overlap, mean, std)
        return lstm_model.predict_proba(out["X"]), {"mean": out["mean"], "std": out["std"]}

    yield from _H2DRing(dev).run(host_batches, prepare, compute)


def _lstm_probs_device(lstm_model, X, batch_size, want_attn, device, autocast=False):
    """`autocast=True`: run where the reference enters `with autocast():` (06_lstm_ode_integration.py:348-351) -- a model built with
    precision="auto" then takes its reduced-precision (bf16 tensor-core) engine exactly there, and its fp32 engine in the callers
    that do not autocast (06:216-234, 08:203-208, 10:226-231).  An explicit precision= on the model always wins."""
    if autocast:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return _lstm_probs_device_impl(lstm_model, X, batch_size, want_attn, device)
    return _lstm_probs_device_impl(lstm_model, X, batch_size, want_attn, device)


def _lstm_probs_device_impl(lstm_model, X, batch_size, want_attn, device):
    """All-window inference returning device tensors; X numpy/CPU tensor (pipelined H2D) or CUDA tensor.

    `batch_size` is the reference callers' argument (512 in predict_batch 06:308, 256 in 08:198, 512 in 10:204).  There it only
    bounds the reference's own memory use; windows are independent, so the result does not depend on it.  Here it is a LOWER
    bound on the pass size: a pass is at least one full wave of the recurrence kernel (`bci_lstm_chunk_windows`, 16 896 windows
    for the bf16 path on a 148-SM B200) -- a drop-in caller that keeps the reference's 512 would otherwise run the GPU at a
    tenth of its rate."""
    dev = _dev(device)
    lstm_model.eval()
    n = len(X)
    wave = ops.lstm_chunk_windows(lstm_model._engine(lstm_model._precision_now()))
    piece = max(int(batch_size or 0), int(wave), 1)
    if n == 0:
        return (torch.empty((0, lstm_model.num_classes), device=dev),
                torch.empty((0, 0), device=dev) if want_attn else None)
    if isinstance(X, torch.Tensor) and X.is_cuda:
        probs = torch.empty((n, lstm_model.num_classes), device=dev, dtype=torch.float32)
        attn = None
        with torch.no_grad():
            for i in range(0, n, piece):
                xb = X[i:i + piece]
                if want_attn:
                    p, a = lstm_model.predict_proba(xb, return_attention=True)
                    if attn is None:
                        attn = torch.empty((n, a.shape[1]), device=dev, dtype=torch.float32)
                    attn[i:i + len(xb)] = a
                else:
                    p = lstm_model.predict_proba(xb)
                probs[i:i + len(xb)] = p
        return probs, attn
    # host input: one pass-sized piece at a time through the double-buffered H2D pipeline (device and pinned staging memory stay
    # bounded by two pieces however long X is)
    outs = list(stream_lstm_probs(lstm_model, (X[i:i + piece] for i in range(0, n, piece)), dev, chunk=piece, want_attn=want_attn))
    probs = outs[0][0] if len(outs) == 1 else torch.cat([o[0] for o in outs])
    attn = None
    if want_attn:
        attn = outs[0][1] if len(outs) == 1 else torch.cat([o[1] for o in outs])
    return probs, attn


class LSTMODEIntegration:
    """Drop-in for 06_lstm_ode_integration.py:183-406."""

    def __init__(self, lstm_model, ode_model, coupling_strength=0.5, device=None, substeps=8):
        self.lstm_model = lstm_model
        self.ode_model = ode_model
        self.coupling_strength = coupling_strength
        self.base_params = ode_model.params.copy()
        self.device = device
        self.substeps = substeps

    def get_lstm_probabilities(self, X):
        """(probs (B,2) [P(open),P(closed)], attention (B,T)) as numpy (06:216-234)."""
        probs, attn = _lstm_probs_device(self.lstm_model, X, max(len(X), 1), True, self.device)
        return probs.cpu().numpy(), attn.cpu().numpy()

    def modulate_ode_rates(self, p_closed, p_open):
        """06:236-264 for one sample (host dict, as the reference returns); the batched path applies the
        same float32 arithmetic inside the ODE kernel."""
        a = np.float32(self.coupling_strength)
        pc, po = np.float32(p_closed), np.float32(p_open)
        prm = self.base_params.copy()
        fat = np.float32(1) + a * pc
        rec = np.float32(1) + a * po
        prm["k_af"] = np.float32(prm["k_af"]) * fat
        prm["k_pf"] = np.float32(prm["k_pf"]) * fat
        prm["k_fa"] = np.float32(prm["k_fa"]) * rec
        prm["k_pa"] = np.float32(prm["k_pa"]) * rec
        return {k: max(0.001, float(v)) for k, v in prm.items()}

    def _solve(self, probs_dev, forecast_steps, y0=None):
        n = probs_dev.shape[0]
        p_open = probs_dev[:, 0].contiguous()
        p_closed = probs_dev[:, 1].contiguous()
        return solve_ensemble(n, p_open=p_open, p_closed=p_closed, base_rates=self.base_params,
                              alpha=self.coupling_strength, y0=y0, y0_mode="given" if y0 is not None else "probs06",
                              coupling=True, style="ref06", mode="rk4", t_end=float(forecast_steps),
                              n_points=int(forecast_steps), substeps=self.substeps, device=self.device)

    def predict_trajectory(self, X, initial_state=None, forecast_steps=10):
        """(trajectory (steps,3) f64, probs (1,2), attention (1,T)) -- 06:266-306."""
        probs, attn = _lstm_probs_device(self.lstm_model, X, max(len(X), 1), True, self.device)
        y0 = None
        if initial_state is not None:
            y0 = torch.tensor(np.asarray(initial_state, dtype=np.float64).reshape(3, 1), dtype=torch.float32)
        traj, _, _ = self._solve(probs[:1], forecast_steps, y0)
        return traj[0].double().cpu().numpy(), probs.cpu().numpy(), attn.cpu().numpy()

    def predict_batch(self, X_batch, forecast_steps=20, batch_size=512, show_progress=True):
        """(trajectories (N,steps,3) f64, probs (N,2) f32, predictions (N,) int) -- 06:308-406."""
        probs, _ = _lstm_probs_device(self.lstm_model, X_batch, batch_size, False, self.device, autocast=True)   # 06:348-351
        traj, final, _ = self._solve(probs, forecast_steps)
        pred, _ = ops.ode_classify(final, True, False)
        return traj.double().cpu().numpy(), probs.cpu().numpy(), pred.cpu().numpy().astype(np.int64)

    def predict_batch_device(self, X_dev, forecast_steps=20, batch_size=None, want_traj=True):
        """Same computation with every tensor left on the device (used by the sharded pipeline)."""
        probs, _ = _lstm_probs_device(self.lstm_model, X_dev, batch_size, False, self.device, autocast=True)
        n = probs.shape[0]
        traj, final, _ = solve_ensemble(n, p_open=probs[:, 0].contiguous(), p_closed=probs[:, 1].contiguous(),
                                        base_rates=self.base_params, alpha=self.coupling_strength, y0_mode="probs06",
                                        coupling=True, style="ref06", mode="rk4", t_end=float(forecast_steps),
                                        n_points=int(forecast_steps), substeps=self.substeps, want_traj=want_traj,
                                        device=self.device)
        pred, cls = ops.ode_classify(final, True, True)
        return traj, probs, final, pred, cls


def get_three_state_probabilities(lstm_model, ode_model, X, batch_size=512, device=None, substeps=8):
    """10:204-290 -> (lstm_probs (N,2) f32, three_state (N,3) f64, predictions (N,))."""
    probs, _ = _lstm_probs_device(lstm_model, X, batch_size, False, device)
    n = probs.shape[0]
    _, final, _ = solve_ensemble(n, p_open=probs[:, 0].contiguous(), p_closed=probs[:, 1].contiguous(),
                                 base_rates=ode_model.params, alpha=0.5, y0_mode="probs06", coupling=True,
                                 style="ref06", mode="rk4", t_end=20.0, n_points=20, substeps=substeps,
                                 want_traj=False, device=device)
    _, cls = ops.ode_classify(final, False, True)
    return probs.cpu().numpy(), final.double().cpu().numpy(), cls.cpu().numpy().astype(np.int64)


# ---- 08_forecasting.py mirrors ----------------------------------------------------------------
def get_lstm_probabilities(lstm_model, X_data, batch_size=256, device=None):
    """08:198-212 -> (N,2) numpy."""
    probs, _ = _lstm_probs_device(lstm_model, X_data, batch_size, False, device)
    return probs.cpu().numpy()


def prob_to_ode_state(prob_closed):
    """08:215-234 for one probability (host scalar helper; float32 arithmetic when given a float32,
    as NumPy >= 2 evaluates the reference).  The batched path derives y0 inside the ODE kernel."""
    p = prob_closed
    wt = np.float32 if isinstance(p, np.float32) else np.float64
    p = wt(p)
    A = wt(1.0) - p
    if p > 0.5:
        F, P = p * wt(0.6), p * wt(0.4)
    else:
        F, P = p * wt(0.3), p * wt(0.3)
    total = A + P + F
    return np.array([A / total, P / total, F / total])


def predict_trajectory(initial_state, params, n_steps, dt=1.0, device=None, substeps=8):
    """08:149-153 -> (n_steps+1, 3) float64: raw rates, no clamp, no renormalisation."""
    y0 = torch.tensor(np.asarray(initial_state, dtype=np.float64).reshape(3, 1), dtype=torch.float32)
    traj, _, _ = solve_ensemble(1, base_rates=params, y0=y0, y0_mode="given", coupling=False, style="ref08",
                                mode="rk4", t_end=float(n_steps) * float(dt), n_points=int(n_steps) + 1,
                                substeps=substeps, f64=True, device=device)
    return traj[0].cpu().numpy()


def _forecast_device(p_closed_dev, ode_params, max_horizon, horizons, device, substeps):
    n = p_closed_dev.shape[0]
    traj, _, _ = solve_ensemble(n, p_closed=p_closed_dev, base_rates=ode_params, y0_mode="pclosed08", coupling=False,
                                style="ref08", mode="rk4", t_end=float(max_horizon), n_points=int(max_horizon) + 1,
                                substeps=substeps, device=device)
    return ops.ode_forecast_readout(traj, horizons)


def multistep_forecast(probs, ode_params, horizons=[5, 10, 20], device=None, substeps=8):
    """08:252-289 -> {h: {'predictions': (N-maxh,), 'actuals': (N-maxh,)}}."""
    dev = _dev(device)
    probs_t = torch.as_tensor(probs, dtype=torch.float32).to(dev)
    max_h = max(horizons)
    m = len(probs_t) - max_h
    results = {h: {"predictions": np.zeros(0), "actuals": np.zeros(0)} for h in horizons}
    if m <= 0:
        return results
    pred = _forecast_device(probs_t[:m, 1].contiguous(), ode_params, max_h, list(horizons), dev, substeps).cpu().numpy()
    pc = probs_t[:, 1].cpu().numpy()
    for j, h in enumerate(horizons):
        results[h]["predictions"] = pred[:, j].astype(np.float64)
        results[h]["actuals"] = pc[h:h + m]
    return results


def rolling_forecast_evaluation(probs, ode_params, window_size=50, horizon=10, device=None, substeps=8):
    """08:346-392 -> pandas DataFrame[window, accuracy, mae]."""
    import pandas as pd
    dev = _dev(device)
    probs_t = torch.as_tensor(probs, dtype=torch.float32).to(dev)
    n = len(probs_t)
    n_windows = (n - window_size - horizon) // window_size
    rows = []
    if n_windows <= 0:
        return pd.DataFrame(rows)
    last = min(n_windows * window_size, n - horizon)
    pred = _forecast_device(probs_t[:last, 1].contiguous(), ode_params, horizon, [horizon], dev, substeps)[:, 0].cpu().numpy()
    pc = probs_t[:, 1].cpu().numpy()
    for w in range(n_windows):
        s, e = w * window_size, min((w + 1) * window_size, n - horizon)
        if e <= s:
            continue
        p, a = pred[s:e], pc[s + horizon:e + horizon]
        rows.append({"window": w, "accuracy": np.mean((p > 0.5) == (a > 0.5)), "mae": np.mean(np.abs(p - a))})
    return pd.DataFrame(rows)
