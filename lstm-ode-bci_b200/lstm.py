"""Host-side mirror of the reference's BiLSTM-attention classifier.

`EnhancedLSTMModel` keeps the reference's constructor signature, `.forward(x, return_attention)`
contract, train/eval switches and -- most importantly -- its state-dict key names and shapes
(04_lstm_model.py:163-204), so a checkpoint written by 04_lstm_model.py:921-933
loads unchanged (`load_state_dict`) and the 06/08/10 scripts can use this class in place of
their own copy.  The module tree below exists only to hold parameters under those names; all
arithmetic is the CUDA path behind `bci::lstm_attn_forward`.
"""
import torch
from torch import nn

from . import _native as N
from . import ops


class _Pool(nn.Module):
    def __init__(self, width):
        super().__init__()
        self.attention = nn.Sequential(nn.Linear(width, width // 2), nn.Tanh(), nn.Linear(width // 2, 1))


class EnhancedLSTMModel(nn.Module):
    """Drop-in for the reference class of the same name (04_lstm_model.py:153-222).

    precision: "fp32" (parity mode, CUDA-core FMA, <=1e-5 on logits/probabilities) or "bf16"
    (tcgen05 tensor-core mode).  Under `torch.autocast("cuda")` -- which the reference enables on
    GPUs (04:486-490, 06:348-351) -- "auto" selects bf16, otherwise fp32.

    `use_attention` / `use_layer_norm` / `bidirectional=False` are the ablation switches of
    09_sensitivity_analysis.py:176-240 (see AblationLSTMModel below); those variants run in fp32 precision.
    """

    def __init__(self, input_size=14, hidden_size=128, num_layers=3, num_classes=2, dropout=0.4,
                 bidirectional=True, num_heads=4, precision="auto", use_attention=True, use_layer_norm=True):
        super().__init__()
        self.hidden_size, self.num_layers, self.bidirectional = hidden_size, num_layers, bool(bidirectional)
        self.use_attention, self.use_layer_norm = bool(use_attention), bool(use_layer_norm)
        self.num_directions = 2 if bidirectional else 1
        self.input_size, self.num_classes, self.dropout_p = input_size, num_classes, dropout
        self.precision = precision
        # training step: "fp32" (parity), "mixed" (tensor-core recurrences with 16-bit operands) or "auto" = mixed exactly where the
        # reference trains in reduced precision, i.e. inside torch.autocast (04:486-490), fp32 elsewhere
        self.train_precision = "auto"
        d = self.num_directions * hidden_size
        self.input_proj = nn.Sequential(nn.Linear(input_size, hidden_size),
                                        nn.LayerNorm(hidden_size) if use_layer_norm else nn.Identity(),
                                        nn.GELU(), nn.Dropout(dropout / 2))
        self.lstm = nn.LSTM(hidden_size, hidden_size, num_layers, batch_first=True,
                            dropout=dropout if num_layers > 1 else 0, bidirectional=self.bidirectional)
        self.layer_norm = nn.LayerNorm(d) if use_layer_norm else nn.Identity()
        self.attention = _Pool(d) if use_attention else None
        self.classifier = nn.Sequential(nn.Linear(d, hidden_size), nn.GELU(), nn.Dropout(dropout),
                                        nn.Linear(hidden_size, hidden_size // 2), nn.GELU(), nn.Dropout(dropout),
                                        nn.Linear(hidden_size // 2, num_classes))
        self._engines = {}       # precision -> handle id (owned by THIS object; never shared with copies)
        self._loaded = {}        # precision -> signature of the parameters at the last load
        self._weights_gen = 0    # bumped by whoever rewrites parameters through raw pointers (train.FusedTrainer)

    @property
    def is_full_model(self):
        return self.bidirectional and self.use_attention and self.use_layer_norm

    # -- engine management -------------------------------------------------------------------
    def _precision_now(self):
        if self.precision == "auto":
            return "bf16" if (torch.is_autocast_enabled("cuda") and self.is_full_model and self.hidden_size in (128, 256)) else "fp32"
        return self.precision

    def _train_precision_now(self):
        if self.train_precision == "auto":
            return "mixed" if (torch.is_autocast_enabled("cuda") and self.hidden_size in (128, 256)) else "fp32"
        return self.train_precision

    def _signature(self):
        return (self._weights_gen,) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def mark_weights_changed(self):
        """Call after parameters were rewritten behind torch's back (raw-pointer optimizer kernels): `data_ptr`/`_version` do not
        change then, so the packed engine copies would otherwise stay one step behind (fp32) or frozen (bf16)."""
        self._weights_gen += 1

    # engine handles are raw library objects: a copy (copy.deepcopy for a best-model snapshot / EMA, pickling) starts without
    # any and packs its own on first use, instead of sharing -- and later double-freeing -- the original's
    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = {} if k in ("_engines", "_loaded") else copy.deepcopy(v, memo)
        return new

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_engines"], st["_loaded"] = {}, {}
        return st

    def _device_of(self, x):
        dev = self.input_proj[0].weight.device
        if x.device != dev:
            raise N.BciError(-1, "input is on %s but the model's parameters are on %s" % (x.device, dev))
        return dev

    def _engine(self, prec):
        if prec not in ("fp32", "bf16"):
            raise N.BciError(-1, "precision must be fp32, bf16 or auto")
        dev = self.input_proj[0].weight.device
        if dev.type != "cuda":
            raise N.BciError(-1, "the model's parameters are on %s; move it with .to('cuda') -- there is no CPU fallback" % dev)
        key = (prec, dev.index if dev.index is not None else torch.cuda.current_device())
        with torch.cuda.device(dev):        # the handle's packed store lives on the parameters' device
            hid = self._engines.get(key)
            if hid is None:
                hid = ops.lstm_create(self.input_size, self.hidden_size, self.num_layers, self.num_classes,
                                      N.PRECISION_BF16 if prec == "bf16" else N.PRECISION_FP32, self.bidirectional,
                                      self.use_attention, self.use_layer_norm)
                self._engines[key] = hid
            sig = self._signature()
            if self._loaded.get(key) != sig:
                ops.lstm_load_weights(hid, {k: v for k, v in self.state_dict().items()})
                self._loaded[key] = sig
        return hid

    def __del__(self):
        try:
            for hid in self._engines.values():
                ops.lstm_destroy(hid)
        except Exception:
            pass

    # -- the reference's forward contract ----------------------------------------------------
    def forward(self, x, return_attention=False):
        if not x.is_cuda:
            raise N.BciError(-1, "EnhancedLSTMModel (bci_b200) runs on CUDA tensors only; move the model and "
                                 "input with .to('cuda') -- there is no CPU fallback")
        x = x.float()
        if torch.is_grad_enabled() and (x.requires_grad or (self.training and any(p.requires_grad for p in self.parameters()))):
            from .train import lstm_attn_autograd
            return lstm_attn_autograd(self, x, return_attention)
        with torch.cuda.device(self._device_of(x)):       # kernels, packed weights and the stream belong to x's device
            hid = self._engine(self._precision_now())
            logits, _probs, attn = ops.lstm_attn_forward(x, hid, bool(return_attention))
        return (logits, attn) if return_attention else logits

    @torch.no_grad()
    def predict_proba(self, x, return_attention=False):
        """logits -> softmax fused in the head kernel: columns [P(open), P(closed)] (06:223,232)."""
        if not x.is_cuda:
            raise N.BciError(-1, "predict_proba runs on CUDA tensors only (there is no CPU fallback)")
        with torch.cuda.device(self._device_of(x)):
            hid = self._engine(self._precision_now())
            _logits, probs, attn = ops.lstm_attn_forward(x.float(), hid, bool(return_attention))
        return (probs, attn) if return_attention else probs


    @torch.no_grad()
    def predict_proba_recordings(self, recordings, seq_len=256, step=128, first_window=0, n_windows=None, return_attention=False):
        """Probabilities of the overlapping windows of normalised recordings read IN PLACE (`bci::lstm_attn_forward_view`).

        recordings: (R, S, C) CUDA tensor, float32 or bfloat16, sample-major -- 02_preprocessing.py's normalised signal
        transposed to (samples, channels).  Window w = r * n_seq + i (n_seq = (S - seq_len) // step + 1, 02:169) is samples
        [i*step, i*step + seq_len) of recording r: the order create_sequences (02:157-180) emits them in, recording after
        recording.  Returns probabilities of windows [first_window, first_window + n_windows) -- default: all of them."""
        if not recordings.is_cuda or recordings.dim() != 3 or recordings.shape[2] != self.input_size:
            raise N.BciError(-1, "recordings must be a CUDA tensor (R, S, %d)" % self.input_size)
        R, S, Cc = (int(v) for v in recordings.shape)
        if S < seq_len:
            raise N.BciError(-1, "recordings of %d samples are shorter than one window (%d)" % (S, seq_len))
        n_seq = (S - seq_len) // step + 1
        total = R * n_seq
        n = total - first_window if n_windows is None else int(n_windows)
        if first_window < 0 or n < 0 or first_window + n > total:
            raise N.BciError(-1, "windows [%d, %d) are outside the %d windows of these recordings" % (first_window, first_window + n, total))
        data = recordings.contiguous()
        with torch.cuda.device(self._device_of(data)):
            hid = self._engine(self._precision_now())
            _logits, probs, attn = ops.lstm_attn_forward_view(data.view(-1), hid, n, int(seq_len), n_seq, int(step) * Cc, S * Cc,
                                                             int(first_window), bool(return_attention))
        return (probs, attn) if return_attention else probs


class AblationLSTMModel(EnhancedLSTMModel):
    """Drop-in for 09_sensitivity_analysis.py:176-240: same constructor (defaults included), `forward(x) -> logits` only.
    The six configurations of run_architecture_ablation (09:330-378) -- full, no attention (mean pooling), unidirectional,
    1 and 2 layers, minimal -- and use_layer_norm=False all map onto switches of the same CUDA op; training goes through
    the same autograd bridge, so quick_train_evaluate (09:265-327) runs unchanged."""

    def __init__(self, input_size=61, hidden_size=256, num_layers=3, num_classes=2, dropout=0.4, bidirectional=True,
                 use_attention=True, use_layer_norm=True, precision="fp32"):
        super().__init__(input_size, hidden_size, num_layers, num_classes, dropout, bidirectional, precision=precision,
                         use_attention=use_attention, use_layer_norm=use_layer_norm)

    def forward(self, x):
        return super().forward(x, return_attention=False)


def from_params(params, precision="fp32", device="cuda", dropout=0.4):
    """Build a CUDA model from a {state-dict key: ndarray/tensor} dict (variant switches inferred from the keys)."""
    H, Cc = params["input_proj.0.weight"].shape
    layers = 0
    while f"lstm.weight_hh_l{layers}" in params:
        layers += 1
    classes = params["classifier.6.weight"].shape[0]
    m = EnhancedLSTMModel(Cc, H, layers, classes, dropout, "lstm.weight_hh_l0_reverse" in params, precision=precision,
                          use_attention="attention.attention.0.weight" in params,
                          use_layer_norm="layer_norm.weight" in params)
    m.load_state_dict({k: torch.as_tensor(v).float() for k, v in params.items()}, strict=True)
    return m.to(device).eval()
