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


def _stream(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


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
        self.device = torch.cuda.current_device()      # the library allocates the packed store on the current device
        self.keepalive = None
        self.layout = []

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
    off, h.layout = 0, []
    for k, v in state.items():
        h.layout.append((k, off, v.numel(), tuple(v.shape)))
        off += v.numel()


PHASES = ("input_proj", "proj_gemm", "recurrence", "pool_head")


TRAIN_MODES = {"fp32": 0, "mixed": 1}   # include/bci_b200.h: BCI_TRAIN_FP32 / BCI_TRAIN_MIXED


def lstm_set_train_mode(hid, mode):
    """Precision of the training step of this handle: "fp32" (parity) or "mixed" (16-bit tensor-core recurrences, single-pass TF32
    GEMMs -- the analogue of the reference's autocast training, 04_lstm_model.py:486-490)."""
    if mode not in TRAIN_MODES:
        raise N.BciError(-1, "train precision must be fp32 or mixed")
    N.check(N.lib().bci_lstm_set_train_mode(_handles[hid].ptr, TRAIN_MODES[mode]))


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
    _same_device(h, x, "x")
    B, T = int(x.shape[0]), int(x.shape[1])
    logits = torch.empty((B, h.cfg.num_classes), device=x.device, dtype=torch.float32)
    probs = torch.empty_like(logits)
    attn = torch.empty((B, T) if want_attn else (0,), device=x.device, dtype=torch.float32)
    if B == 0:
        return logits, probs, attn
    nbytes = lstm_workspace_bytes(handle, B, T, 0)
    ws = torch.empty((nbytes,), device=x.device, dtype=torch.uint8)
    N.check(N.lib().bci_lstm_forward(h.ptr, _ptr(x), B, T, 0, 0.0, 0, _ptr(logits), _ptr(probs),
                                     _ptr(attn) if want_attn else C.c_void_p(0), _ptr(ws), nbytes, _stream(x.device)))
    return logits, probs, attn


@lstm_attn_forward.register_fake
def _(x, handle, want_attn):
    h = _handles[handle]
    B, T = x.shape[0], x.shape[1]
    lg = x.new_empty((B, h.cfg.num_classes))
    return lg, x.new_empty((B, h.cfg.num_classes)), x.new_empty((B, T) if want_attn else (0,))


def _same_device(h, t, name):
    if t.device.index != h.device:
        raise N.BciError(-1, "%s is on cuda:%s but the engine was built on cuda:%d" % (name, t.device.index, h.device))


@torch.library.custom_op("bci::lstm_attn_forward_view", mutates_args=())
def lstm_attn_forward_view(data: torch.Tensor, handle: int, batch: int, seq_len: int, windows_per_run: int, window_stride: int,
                           run_stride: int, first_window: int, want_attn: bool) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """The same forward over windows read IN PLACE from `data` (bci_lstm_forward_view): a flat fp32 or bf16 CUDA tensor holding
    e.g. (R, S, C) recordings whose 50 %-overlapping windows (02_preprocessing.py:157-180) start every `window_stride` elements.
    Window b covers elements [off, off + seq_len*C), off = (w // windows_per_run) * run_stride + (w % windows_per_run) * window_stride
    with w = first_window + b (windows_per_run = 0: w * window_stride)."""
    h = _handles[handle]
    if not (isinstance(data, torch.Tensor) and data.is_cuda and data.is_contiguous()):
        raise N.BciError(-1, "data must be a contiguous CUDA tensor (there is no CPU fallback)")
    if data.dtype not in (torch.float32, torch.bfloat16):
        raise N.BciError(-1, "data must be float32 or bfloat16, got %s" % data.dtype)
    _same_device(h, data, "data")
    B, T, Cc = int(batch), int(seq_len), h.cfg.input_size
    if B > 0:
        last = first_window + B - 1
        off = (last // windows_per_run) * run_stride + (last % windows_per_run) * window_stride if windows_per_run > 0 \
            else last * window_stride
        if window_stride < 1 or first_window < 0 or off + T * Cc > data.numel():
            raise N.BciError(-1, "view of %d windows x %d x %d reaches element %d of a %d-element tensor"
                             % (B, T, Cc, off + T * Cc, data.numel()))
    logits = torch.empty((B, h.cfg.num_classes), device=data.device, dtype=torch.float32)
    probs = torch.empty_like(logits)
    attn = torch.empty((B, T) if want_attn else (0,), device=data.device, dtype=torch.float32)
    if B == 0:
        return logits, probs, attn
    nbytes = lstm_workspace_bytes(handle, B, T, 0)
    ws = torch.empty((nbytes,), device=data.device, dtype=torch.uint8)
    view = N.LstmInput(data.data_ptr(), N.IN_BF16 if data.dtype == torch.bfloat16 else N.IN_F32, int(windows_per_run),
                       int(window_stride), int(run_stride), int(first_window))
    N.check(N.lib().bci_lstm_forward_view(h.ptr, C.byref(view), B, T, _ptr(logits), _ptr(probs),
                                          _ptr(attn) if want_attn else C.c_void_p(0), _ptr(ws), nbytes, _stream(data.device)))
    return logits, probs, attn


@lstm_attn_forward_view.register_fake
def _(data, handle, batch, seq_len, windows_per_run, window_stride, run_stride, first_window, want_attn):
    h = _handles[handle]
    lg = data.new_empty((batch, h.cfg.num_classes), dtype=torch.float32)
    return lg, torch.empty_like(lg), data.new_empty((batch, seq_len) if want_attn else (0,), dtype=torch.float32)


# ------------------------------------------------------------------------------------------
# training step: forward with saved activations, BPTT (the autograd bridge in train.py calls these two ops)
# ------------------------------------------------------------------------------------------
@torch.library.custom_op("bci::lstm_attn_forward_train", mutates_args=())
def lstm_attn_forward_train(x: torch.Tensor, handle: int, dropout: float, seed: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """EnhancedLSTMModel.forward in .train() mode (04_lstm_model.py:206-222 with the dropout sites of 04:177,186,199,202):
    x (B,T,C) fp32 -> logits (B,classes), attention (B,T), and the opaque workspace holding the activations BPTT needs."""
    h = _handles[handle]
    x = _need_cuda(x, "x")
    _same_device(h, x, "x")
    B, T = int(x.shape[0]), int(x.shape[1])
    logits = torch.empty((B, h.cfg.num_classes), device=x.device, dtype=torch.float32)
    attn = torch.empty((B, T), device=x.device, dtype=torch.float32)
    nbytes = lstm_workspace_bytes(handle, B, T, 1)
    ws = torch.empty((nbytes,), device=x.device, dtype=torch.uint8)
    N.check(N.lib().bci_lstm_forward(h.ptr, _ptr(x), B, T, 1, float(dropout), int(seed), _ptr(logits), C.c_void_p(0), _ptr(attn),
                                     _ptr(ws), nbytes, _stream(x.device)))
    return logits, attn, ws


@lstm_attn_forward_train.register_fake
def _(x, handle, dropout, seed):
    h = _handles[handle]
    B, T = x.shape[0], x.shape[1]
    return (x.new_empty((B, h.cfg.num_classes)), x.new_empty((B, T)),
            x.new_empty((lstm_workspace_bytes(handle, int(B), int(T), 1),), dtype=torch.uint8))


@torch.library.custom_op("bci::lstm_attn_backward", mutates_args=("workspace",))
def lstm_attn_backward(x: torch.Tensor, dlogits: torch.Tensor, workspace: torch.Tensor, handle: int,
                       need_dx: bool) -> tuple[torch.Tensor, torch.Tensor]:
    """loss.backward() through the model (04_lstm_model.py:490-494; 07_explainability.py:242-258 for dx): BPTT from
    dlogits (B,classes) on the workspace of a `bci::lstm_attn_forward_train` call.  Returns (dx (B,T,C) or empty,
    flat fp32 gradient bucket in state-dict order: `param_layout(handle)` gives each parameter's (key, offset, numel, shape))."""
    h = _handles[handle]
    x = _need_cuda(x, "x")
    _same_device(h, x, "x")
    B, T = int(x.shape[0]), int(x.shape[1])
    lay = param_layout(handle)
    flat = torch.empty((lay[-1][1] + lay[-1][2],), device=x.device, dtype=torch.float32)
    grads = {k: flat[o:o + n] for k, o, n, _ in lay}
    gs = fill_pointer_struct(N.LstmGrads(), grads, h.cfg.num_layers)
    dx = torch.empty_like(x) if need_dx else torch.empty((0,), device=x.device, dtype=torch.float32)
    dl = _need_cuda(dlogits.float(), "dlogits")
    N.check(N.lib().bci_lstm_backward(h.ptr, _ptr(x), _ptr(dl), B, T, _ptr(dx) if need_dx else C.c_void_p(0), C.byref(gs),
                                      _ptr(workspace), workspace.numel(), _stream(x.device)))
    return dx, flat


@lstm_attn_backward.register_fake
def _(x, dlogits, workspace, handle, need_dx):
    lay = param_layout(handle)
    return (torch.empty_like(x) if need_dx else x.new_empty((0,))), x.new_empty((lay[-1][1] + lay[-1][2],))


def param_layout(hid):
    """[(state-dict key, offset, numel, shape)] of the parameters last loaded into the handle, in state-dict order."""
    lay = _handles[hid].layout
    if not lay:
        raise N.BciError(-4, "no weights loaded into this engine yet")
    return lay


def ce_loss_grad(logits, labels, class_weight=None, loss_scale=1.0):
    """criterion(outputs, y) * loss_scale and its gradient wrt the logits (04:456-458,486-494): (loss (1,), dlogits (B,classes))."""
    lg = _need_cuda(logits, "logits")
    y = _need_cuda(labels, "labels", torch.int64)
    cw = _need_cuda(class_weight, "class_weight") if class_weight is not None else None
    loss = torch.empty((1,), device=lg.device, dtype=torch.float32)
    dl = torch.empty_like(lg)
    N.check(N.lib().bci_ce_loss_grad(_ptr(lg), _ptr(y), _ptr(cw), int(lg.shape[0]), int(lg.shape[1]), float(loss_scale), _ptr(loss),
                                     _ptr(dl), _stream(lg.device)))
    return loss, dl


def grad_accumulate(acc, g, first):
    N.check(N.lib().bci_grad_accumulate(_ptr(acc), _ptr(g), acc.numel(), int(bool(first)), _stream(acc.device)))


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


def permute_channels(x, perm, channel, row0, rows, bf16_out=False):
    """Rows [row0, row0 + rows) of the V x n stack of channel-permuted copies of x (n,T,C) fp32 (bci_permute_channels;
    07_explainability.py:336-339): perm (V*n) int32, channel (V) int32 (< 0: unpermuted copy).  -> (rows,T,C) fp32 / bf16."""
    x = _need_cuda(x, "x")
    perm = _need_cuda(perm, "perm", torch.int32)
    channel = _need_cuda(channel, "channel", torch.int32)
    n, T, Cc = (int(v) for v in x.shape)
    if perm.numel() != channel.numel() * n or row0 < 0 or row0 + rows > perm.numel():
        raise N.BciError(-1, "permute_channels: perm must hold %d x %d indices and rows [%d, %d) must lie inside them"
                         % (channel.numel(), n, row0, row0 + rows))
    out = torch.empty((int(rows), T, Cc), device=x.device, dtype=torch.bfloat16 if bf16_out else torch.float32)
    with torch.cuda.device(x.device):
        N.check(N.lib().bci_permute_channels(_ptr(x), n, T, Cc, _ptr(perm), _ptr(channel), int(row0), int(rows),
                                             N.IN_BF16 if bf16_out else N.IN_F32, _ptr(out), _stream(x.device)))
    return out


def fp32_peak_probe():
    v = C.c_double(0.0)
    N.check(N.lib().bci_fp32_peak_probe(C.byref(v), _stream()))
    return v.value


def fp64_peak_probe():
    v = C.c_double(0.0)
    N.check(N.lib().bci_fp64_peak_probe(C.byref(v), _stream()))
    return v.value
