"""CPU oracle for the BiLSTM + attention-pooling forward (TEST INFRASTRUCTURE ONLY).

A plain-numpy restatement of the reference's `EnhancedLSTMModel.forward`
(/root/reference/04_lstm_model.py:206-222) in eval mode, from the module definitions at
04_lstm_model.py:112-128 (Attention) and 04_lstm_model.py:163-204 (layers), plus the
`torch.nn.LSTM` cell equations the reference delegates to (torch, version unpinned by the
reference: requirements.txt:5 `torch>=2.0.0`; 2.11.0 in this image):

    i,f,g,o = split(W_ih x_t + b_ih + W_hh h_{t-1} + b_hh)      gate row order i,f,g,o
    c_t = sigmoid(f) c_{t-1} + sigmoid(i) tanh(g);  h_t = sigmoid(o) tanh(c_t);  h_0 = c_0 = 0

Parity pin: the reference has no tests or golden vectors for this path, so
the oracle is pinned against outputs of the reference itself, generated in the build
container by tests/golden/make_golden.py and committed under tests/golden/.
tests/test_oracle_golden.py checks this file against those vectors.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import math
import numpy as np

try:  # erf for exact (erf-based) GELU; scipy is in the image, fall back to math.erf
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _gelu(x):
    # nn.GELU() default = exact erf form (04_lstm_model.py:176,198,201)
    return 0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))


def _layer_norm(x, w, b, eps=1e-5):
    # nn.LayerNorm: biased variance over the last dim, eps inside the sqrt
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * w + b


def input_projection(p, x):
    """04_lstm_model.py:173-178,208 (eval: dropout is identity)."""
    z = x @ p["input_proj.0.weight"].T + p["input_proj.0.bias"]
    if "input_proj.1.weight" in p:  # nn.Identity in the use_layer_norm=False ablation (09:191)
        z = _layer_norm(z, p["input_proj.1.weight"], p["input_proj.1.bias"])
    return _gelu(z)


def lstm_direction(inp, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of one nn.LSTM layer.  inp (B,T,K) -> (B,T,H); the reverse direction
    walks t = T-1..0 and stores h_t at index t (04_lstm_model.py:181-188)."""
    B, T, _ = inp.shape
    H = w_hh.shape[1]
    h = np.zeros((B, H), dtype=inp.dtype)
    c = np.zeros((B, H), dtype=inp.dtype)
    out = np.empty((B, T, H), dtype=inp.dtype)
    bias = b_ih + b_hh
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        g = inp[:, t, :] @ w_ih.T + h @ w_hh.T + bias
        i = _sigmoid(g[:, 0 * H:1 * H])
        f = _sigmoid(g[:, 1 * H:2 * H])
        gg = np.tanh(g[:, 2 * H:3 * H])
        o = _sigmoid(g[:, 3 * H:4 * H])
        c = f * c + i * gg
        h = o * np.tanh(c)
        out[:, t, :] = h
    return out


def lstm_stack(p, z, layers, bidirectional=True):
    inp = z
    for l in range(layers):
        outs = []
        for suf, rev in (("", False),) + ((("_reverse", True),) if bidirectional else ()):
            outs.append(lstm_direction(inp,
                                       p[f"lstm.weight_ih_l{l}{suf}"], p[f"lstm.weight_hh_l{l}{suf}"],
                                       p[f"lstm.bias_ih_l{l}{suf}"], p[f"lstm.bias_hh_l{l}{suf}"], rev))
        inp = np.concatenate(outs, axis=-1)  # [fwd | bwd]
    return inp


def attention_pool(p, y):
    """04_lstm_model.py:112-128: scores -> softmax over T (dim=1) -> weighted sum."""
    u = np.tanh(y @ p["attention.attention.0.weight"].T + p["attention.attention.0.bias"])
    s = u @ p["attention.attention.2.weight"].T + p["attention.attention.2.bias"]  # (B,T,1)
    s = s[..., 0]
    s = s - s.max(axis=1, keepdims=True)
    e = np.exp(s)
    a = e / e.sum(axis=1, keepdims=True)
    ctx = (a[..., None] * y).sum(axis=1)
    return ctx, a


def classifier(p, ctx):
    """04_lstm_model.py:196-204,218 (eval)."""
    h = _gelu(ctx @ p["classifier.0.weight"].T + p["classifier.0.bias"])
    h = _gelu(h @ p["classifier.3.weight"].T + p["classifier.3.bias"])
    return h @ p["classifier.6.weight"].T + p["classifier.6.bias"]


def softmax_probs(logits):
    """Callers' softmax(dim=1): column 0 = P(open), 1 = P(closed) (06:223,232)."""
    z = logits - logits.max(axis=1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=1, keepdims=True)


def infer_config(p):
    H = p["input_proj.0.weight"].shape[0]
    layers = 0
    while f"lstm.weight_hh_l{layers}" in p:
        layers += 1
    bidir = "lstm.weight_hh_l0_reverse" in p
    return H, layers, bidir


def forward(params, x, dtype=np.float64, return_intermediates=False):
    """Full eval-mode forward.  x (B,T,C).  Computes in `dtype` (float64 default: the
    oracle is then ~1e-7 from the reference's fp32 path, dominated by the reference's own
    rounding).  Returns logits (B,classes), attention (B,T)."""
    p = {k: np.asarray(v, dtype=dtype) for k, v in params.items()}
    x = np.asarray(x, dtype=dtype)
    _, layers, bidir = infer_config(p)
    z = input_projection(p, x)
    out = lstm_stack(p, z, layers, bidir)
    # ablation variants (09_sensitivity_analysis.py:176-240): Identity instead of LayerNorm (09:210), mean over time
    # instead of attention pooling (09:229-234)
    y = _layer_norm(out, p["layer_norm.weight"], p["layer_norm.bias"]) if "layer_norm.weight" in p else out
    if "attention.attention.0.weight" in p:
        ctx, attn = attention_pool(p, y)
    else:
        ctx, attn = y.mean(axis=1), np.full(y.shape[:2], 1.0 / y.shape[1], dtype=y.dtype)
    logits = classifier(p, ctx)
    if return_intermediates:
        return logits, attn, {"z": z, "lstm_out": out, "y": y, "ctx": ctx}
    return logits, attn
