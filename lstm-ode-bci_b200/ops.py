"""torch.library custom ops over the C ABI (`bci::lstm_attn_forward`, `bci::ode_ensemble`, ...).

PyTorch is plumbing here: it owns device memory (caching allocator -- callers of the
reference free with `del` + `torch.cuda.empty_cache()`, 06:360-362, so nothing is cached per
call) and the current stream; all arithmetic happens in libbci_b200.so.
"""
import ctypes as C
import threading

import torch

from . import _native as N

_handles = {}
_handles_lock = threading.Lock()
_next_handle = [1]


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _need_cuda(t, name, dtype=torch.float32):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise N.BciError(-1, "%s must be a CUDA tensor (there is no CPU fallback)" % name)
    if t.dtype != dtype:
        raise N.BciError(-1, "%s must be %s, got %s" % (name, dtype, t.dtype))
    return t.contiguous()


# ------------------------------------------------------------------------------------------
# LSTM handle registry (custom ops cannot take opaque pointers; they take an int id)
# ------------------------------------------------------------------------------------------
class _Handle:
    def __init__(self, cfg):
        self.cfg = cfg
        self.ptr = C.c_void_p(0)
        N.check(N.lib().bci_lstm_create(C.byref(cfg), C.byref(self.ptr)))
        self.keepalive = None

    def close(self):
        if self.ptr:
            N.lib().bci_lstm_destroy(self.ptr)
            self.ptr = C.c_void_p(0)


def lstm_create(input_size, hidden_size, num_layers, num_classes, precision, bidirectional=True, use_attention=True,
                use_layer_norm=True):
    cfg = N.LstmConfig(int(input_size), int(hidden_size), int(num_layers), int(num_classes), int(bool(bidirectional)),
                       int(precision), int(bool(use_attention)), int(bool(use_layer_norm)))
    h = _Handle(cfg)
    with _handles_lock:
        hid = _next_handle[0]
        _next_handle[0] += 1
        _handles[hid] = h
    return hid


def lstm_destroy(hid):
    with _handles_lock:
        h = _handles.pop(hid, None)
    if h is not None:
        h.close()


_KEYMAP = [
    ("input_proj_w", "input_proj.0.weight"), ("input_proj_b", "input_proj.0.bias"),
    ("input_ln_w", "input_proj.1.weight"), ("input_ln_b", "input_proj.1.bias"),
    ("ln_w", "layer_norm.weight"), ("ln_b", "layer_norm.bias"),
    ("attn_w1", "attention.attention.0.weight"), ("attn_b1", "attention.attention.0.bias"),
    ("attn_w2", "attention.attention.2.weight"), ("attn_b2", "attention.attention.2.bias"),
    ("cls_w0", "classifier.0.weight"), ("cls_b0", "classifier.0.bias"),
    ("cls_w3", "classifier.3.weight"), ("cls_b3", "classifier.3.bias"),
    ("cls_w6", "classifier.6.weight"), ("cls_b6", "classifier.6.bias"),
]


def fill_pointer_struct(struct, tensors, num_layers):
    """tensors: {state-dict key: contiguous fp32 CUDA tensor} (SURVEY.md §8 a1 names).  Keys of modules an ablation
    variant does not have (LayerNorms, attention, the reverse direction: 09:176-240) are left NULL."""
    for field, key in _KEYMAP:
        setattr(struct, field, tensors[key].data_ptr() if key in tensors else None)
    for l in range(num_layers):
        for d, suf in enumerate(("", "_reverse")):
            if f"lstm.weight_ih_l{l}{suf}" not in tensors:
                continue
            struct.w_ih[l][d] = tensors[f"lstm.weight_ih_l{l}{suf}"].data_ptr()
            struct.w_hh[l][d] = tensors[f"lstm.weight_hh_l{l}{suf}"].data_ptr()
            struct.b_ih[l][d] = tensors[f"lstm.bias_ih_l{l}{suf}"].data_ptr()
            struct.b_hh[l][d] = tensors[f"lstm.bias_hh_l{l}{suf}"].data_ptr()
    return struct


def lstm_load_weights(hid, state):
    h = _handles[hid]
    tensors = {k: _need_cuda(v.detach(), k) for k, v in state.items()}
    w = fill_pointer_struct(N.LstmWeights(), tensors, h.cfg.num_layers)
    N.check(N.lib().bci_lstm_load_weights(h.ptr, C.byref(w), _stream()))
    h.keepalive = tensors  # backward reads the raw weights; keep them alive with the handle


PHASES = ("input_proj", "proj_gemm", "recurrence", "pool_head")


def lstm_set_profiling(hid, enable):
    N.check(N.lib().bci_lstm_set_profiling(_handles[hid].ptr, int(bool(enable))))


def lstm_get_profile(hid):
    """{phase: (milliseconds, launches)} accumulated since the last call (synchronises)."""
    ms = (C.c_float * 4)()
    ln = (C.c_int32 * 4)()
    N.check(N.lib().bci_lstm_get_profile(_handles[hid].ptr, ms, ln))
    return {PHASES[i]: (float(ms[i]), int(ln[i])) for i in range(4)}


def launch_count():
    return int(N.lib().bci_launch_count())


def lstm_chunk_windows(hid):
    """Windows per internal pass of the inference forward (one full wave of the recurrence kernel on this device)."""
    n = C.c_int32(0)
    N.check(N.lib().bci_lstm_chunk_windows(_handles[hid].ptr, C.byref(n)))
    return n.value


def lstm_workspace_bytes(hid, batch, seq_len, train):
    h = _handles[hid]
    n = C.c_size_t(0)
    N.check(N.lib().bci_lstm_workspace_bytes(h.ptr, int(batch), int(seq_len), int(train), C.byref(n)))
    return n.value


@torch.library.custom_op("bci::lstm_attn_forward", mutates_args=())
def lstm_attn_forward(x: torch.Tensor, handle: int, want_attn: bool) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """EnhancedLSTMModel.forward + softmax (04_lstm_model.py:206-222; 06:232).
    x (B,T,C) fp32 CUDA -> logits (B,classes), probs (B,classes), attention (B,T) [empty if not wanted]."""
    h = _handles[handle]
    x = _need_cuda(x, "x")
    if x.dim() != 3 or x.shape[2] != h.cfg.input_size:
        raise N.BciError(-1, "x must be (B,T,%d), got %s" % (h.cfg.input_size, tuple(x.shape)))
    B, T = int(x.shape[0]), int(x.shape[1])
    logits = torch.empty((B, h.cfg.num_classes), device=x.device, dtype=torch.float32)
    probs = torch.empty_like(logits)
    attn = torch.empty((B, T) if want_attn else (0,), device=x.device, dtype=torch.float32)
    if B == 0:
        return logits, probs, attn
    nbytes = lstm_workspace_bytes(handle, B, T, 0)
    ws = torch.empty((nbytes,), device=x.device, dtype=torch.uint8)
    N.check(N.lib().bci_lstm_forward(h.ptr, _ptr(x), B, T, 0, 0.0, 0, _ptr(logits), _ptr(probs),
                                     _ptr(attn) if want_attn else C.c_void_p(0), _ptr(ws), nbytes, _stream()))
    return logits, probs, attn


@lstm_attn_forward.register_fake
def _(x, handle, want_attn):
    h = _handles[handle]
    B, T = x.shape[0], x.shape[1]
    lg = x.new_empty((B, h.cfg.num_classes))
    return lg, x.new_empty((B, h.cfg.num_classes)), x.new_empty((B, T) if want_attn else (0,))


# ------------------------------------------------------------------------------------------
# ODE ensemble
# ------------------------------------------------------------------------------------------
@torch.library.custom_op("bci::ode_ensemble", mutates_args=())
def ode_ensemble(p_open: torch.Tensor | None, p_closed: torch.Tensor | None, rates: torch.Tensor | None,
                 alpha_arr: torch.Tensor | None, y0: torch.Tensor | None, base_rates: list[float], alpha: float,
                 n: int, mode: int, style: int, y0_mode: int, coupling: bool, t_end: float, n_points: int,
                 substeps: int, rtol: float, atol: float, want_traj: bool, want_steps: bool,
                 f64: bool, device_index: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """One launch over N independent trajectories: coupling (06:236-264) + initial state
    (06:377-382 | 08:215-234 | given) + RK4/RK45 solve (06:174-180 | 08:149-153).
    Returns traj (N,n_points,3) [empty if not wanted], final_state (N,3), n_steps (N) [empty if not wanted]."""
    dev = torch.device("cuda", device_index)
    dt = torch.float64 if f64 else torch.float32
    a = N.OdeArgs()
    a.mode, a.style, a.y0_mode, a.coupling, a.n = mode, style, y0_mode, int(coupling), n
    for i in range(6):
        a.base_rates[i] = float(base_rates[i])
    a.alpha = float(alpha)
    keep = []

    def dev_in(t, name, numel):
        if t is None:
            return None
        t = _need_cuda(t, name)
        if t.numel() != numel:
            raise N.BciError(-1, "%s must have %d elements, got %d" % (name, numel, t.numel()))
        keep.append(t)
        return t.data_ptr()

    a.rates = dev_in(rates, "rates", 6 * n)
    a.alpha_arr = dev_in(alpha_arr, "alpha_arr", n)
    a.p_open = dev_in(p_open, "p_open", n)
    a.p_closed = dev_in(p_closed, "p_closed", n)
    a.y0 = dev_in(y0, "y0", 3 * n)
    a.t_end, a.n_points, a.substeps, a.rtol, a.atol = float(t_end), int(n_points), int(substeps), float(rtol), float(atol)
    a.out_dtype = N.OUT_F64 if f64 else N.OUT_F32
    traj = torch.empty((n, n_points, 3) if want_traj else (0,), device=dev, dtype=dt)
    final = torch.empty((n, 3), device=dev, dtype=dt)
    steps = torch.empty((n,) if want_steps else (0,), device=dev, dtype=torch.int32)
    a.traj = traj.data_ptr() if want_traj else None
    a.final_state = final.data_ptr()
    a.n_steps = steps.data_ptr() if want_steps else None
    with torch.cuda.device(dev):
        N.check(N.lib().bci_ode_solve(C.byref(a), _stream()))
    return traj, final, steps


@ode_ensemble.register_fake
def _(p_open, p_closed, rates, alpha_arr, y0, base_rates, alpha, n, mode, style, y0_mode, coupling, t_end, n_points,
      substeps, rtol, atol, want_traj, want_steps, f64, device_index):
    dev = torch.device("cuda", device_index)
    dt = torch.float64 if f64 else torch.float32
    return (torch.empty((n, n_points, 3) if want_traj else (0,), device=dev, dtype=dt),
            torch.empty((n, 3), device=dev, dtype=dt),
            torch.empty((n,) if want_steps else (0,), device=dev, dtype=torch.int32))


def ode_classify(final_state, want_pred06=True, want_cls10=True):
    fs = _need_cuda(final_state, "final_state")
    n = fs.shape[0]
    p = torch.empty((n,), device=fs.device, dtype=torch.int32) if want_pred06 else None
    c = torch.empty((n,), device=fs.device, dtype=torch.int32) if want_cls10 else None
    N.check(N.lib().bci_ode_classify(_ptr(fs), n, _ptr(p), _ptr(c), _stream()))
    return p, c


def ode_forecast_readout(traj, horizons):
    tr = _need_cuda(traj, "traj")
    n, n_points = int(tr.shape[0]), int(tr.shape[1])
    hz = (C.c_int32 * len(horizons))(*[int(h) for h in horizons])
    out = torch.empty((n, len(horizons)), device=tr.device, dtype=torch.float32)
    N.check(N.lib().bci_ode_forecast_readout(_ptr(tr), n, n_points, hz, len(horizons), _ptr(out), _stream()))
    return out


def fp32_peak_probe():
    v = C.c_double(0.0)
    N.check(N.lib().bci_fp32_peak_probe(C.byref(v), _stream()))
    return v.value


def fp64_peak_probe():
    v = C.c_double(0.0)
    N.check(N.lib().bci_fp64_peak_probe(C.byref(v), _stream()))
    return v.value
