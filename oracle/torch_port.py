"""torch-CPU port of the reference model and training step (TEST INFRASTRUCTURE ONLY).

Why a second LSTM oracle next to lstm_oracle.py: the reference's arithmetic *is* torch.nn
(04_lstm_model.py:172-204 builds nn.Linear/nn.LayerNorm/nn.GELU/nn.LSTM; no kernels of its
own), so the faithful CPU baseline -- what a user of the reference runs on the host today --
is torch's own CPU path (oneDNN/MKL, all host threads), and torch autograd over that path is
the gradient oracle for the training step (04_lstm_model.py:482-507).  This module
rebuilds that module tree with the reference's state-dict key names so
weights move between the reference, this port and the CUDA path unchanged.

Pinned by tests/test_oracle_golden.py against tests/golden/ vectors produced by the live
reference (tests/golden/make_golden.py).  Only tests/, smoke() and bench.py's cpu legs may
import it; the product path never does.
"""
import numpy as np
import torch
from torch import nn


def _mlp(sizes, act, drops):
    """Sequential of Linear/act/Dropout triples; keeps the reference's child indices
    (classifier.0/.3/.6, 04_lstm_model.py:196-204)."""
    mods = []
    for i in range(len(sizes) - 1):
        mods.append(nn.Linear(sizes[i], sizes[i + 1]))
        if i < len(sizes) - 2:
            mods += [act(), nn.Dropout(drops)]
    return nn.Sequential(*mods)


class _Pool(nn.Module):
    """Additive attention pooling, child name `attention` with Linear at .0 and .2
    (04_lstm_model.py:112-128)."""

    def __init__(self, width):
        super().__init__()
        self.attention = nn.Sequential(nn.Linear(width, width // 2), nn.Tanh(), nn.Linear(width // 2, 1))

    def forward(self, y):
        w = torch.softmax(self.attention(y), dim=1)
        return (w * y).sum(dim=1), w.squeeze(-1)


class BiLSTMAttnPort(nn.Module):
    """EnhancedLSTMModel (04:153-222); with use_attention / use_layer_norm / bidirectional switched off it is
    AblationLSTMModel (09_sensitivity_analysis.py:176-240)."""

    def __init__(self, input_size=61, hidden_size=128, num_layers=3, num_classes=2,
                 dropout=0.4, bidirectional=True, use_attention=True, use_layer_norm=True):
        super().__init__()
        d = 2 if bidirectional else 1
        self.input_proj = nn.Sequential(nn.Linear(input_size, hidden_size),
                                        nn.LayerNorm(hidden_size) if use_layer_norm else nn.Identity(),
                                        nn.GELU(), nn.Dropout(dropout / 2))
        self.lstm = nn.LSTM(hidden_size, hidden_size, num_layers, batch_first=True,
                            dropout=dropout if num_layers > 1 else 0.0, bidirectional=bidirectional)
        self.layer_norm = nn.LayerNorm(d * hidden_size) if use_layer_norm else nn.Identity()
        self.attention = _Pool(d * hidden_size) if use_attention else None
        self.classifier = _mlp([d * hidden_size, hidden_size, hidden_size // 2, num_classes], nn.GELU, dropout)

    def forward(self, x, return_attention=False):
        seq, _ = self.lstm(self.input_proj(x))
        y = self.layer_norm(seq)
        if self.attention is not None:
            ctx, attn = self.attention(y)
        else:
            ctx, attn = torch.mean(y, dim=1), torch.full(y.shape[:2], 1.0 / y.shape[1])
        logits = self.classifier(ctx)
        return (logits, attn) if return_attention else logits


def build_port(params, dropout=0.4):
    """Instantiate the port from a {key: ndarray} parameter dict (fp32)."""
    H = params["input_proj.0.weight"].shape[0]
    C = params["input_proj.0.weight"].shape[1]
    layers = 0
    while f"lstm.weight_hh_l{layers}" in params:
        layers += 1
    bidir = "lstm.weight_hh_l0_reverse" in params
    classes = params["classifier.6.weight"].shape[0]
    m = BiLSTMAttnPort(C, H, layers, classes, dropout, bidir, "attention.attention.0.weight" in params,
                       "layer_norm.weight" in params)
    m.load_state_dict({k: torch.from_numpy(np.array(v, dtype=np.float32)) for k, v in params.items()}, strict=True)
    return m


def forward_probs(model, x, batch_size=512):
    """Eval forward in the reference's batching (06_lstm_ode_integration.py:339-357, fp32 CPU)."""
    model.eval()
    probs, attn = [], []
    with torch.no_grad():
        for i in range(0, len(x), batch_size):
            xb = torch.as_tensor(x[i:i + batch_size], dtype=torch.float32)
            lg, a = model(xb, return_attention=True)
            probs.append(torch.softmax(lg, dim=1).numpy())
            attn.append(a.numpy())
    return np.concatenate(probs), np.concatenate(attn)


def loss_and_grads(model, x, y, class_weight=None, train_mode_no_dropout=True):
    """Weighted cross-entropy + autograd gradients (04_lstm_model.py:486-494 without AMP/accum).
    Dropout must be zero for gradient parity (SURVEY.md §7 'Train-mode semantics'): callers
    build the port with dropout=0.0; train() is needed for cuDNN-free CPU LSTM backward anyway."""
    if train_mode_no_dropout:
        model.train()
    model.zero_grad(set_to_none=True)
    xb = torch.as_tensor(x, dtype=torch.float32).requires_grad_(True)
    yb = torch.as_tensor(y, dtype=torch.long)
    w = None if class_weight is None else torch.as_tensor(class_weight, dtype=torch.float32)
    logits = model(xb)
    loss = nn.functional.cross_entropy(logits, yb, weight=w)
    loss.backward()
    grads = {k: p.grad.detach().numpy().copy() for k, p in model.named_parameters()}
    return float(loss.detach()), grads, xb.grad.detach().numpy().copy(), logits.detach().numpy().copy()
