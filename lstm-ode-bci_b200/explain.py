"""SURVEY.md §8 f rank 1: gradient-based attribution on the B200 backward pass.

`compute_channel_importance` mirrors 07_explainability.py:203-284 (same arguments, same sampling call, same
DataFrame).  The reference back-propagates outputs[i, pred_i] once PER SAMPLE with retain_graph=True -- B full backward
passes per batch.  Windows are independent, so d outputs[i, pred_i] / d x[i] for every i of a batch is exactly row i of ONE
backward pass with dlogits[i, pred_i] = 1: `input_gradients` does that (B x fewer BPTT launches).  The per-sample pattern
of the reference also works unchanged against lstm.EnhancedLSTMModel (autograd bridge keeps the saved forward)."""
import numpy as np
import torch

from . import _native as N
from . import ops
from .train import lstm_attn_autograd


def input_gradients(lstm_model, X, batch_size=32, device=None, seed=None):
    """d logit[pred] / d x for every window: (N,T,C) float32 numpy.  Train-mode semantics as in the reference
    (dropout active iff lstm_model.training and dropout_p > 0)."""
    dev = device if device is not None else next(lstm_model.parameters()).device
    out = np.empty(tuple(X.shape), dtype=np.float32)
    for i in range(0, len(X), batch_size):
        xb = torch.as_tensor(X[i:i + batch_size], dtype=torch.float32, device=dev).requires_grad_(True)
        with torch.enable_grad():
            logits = lstm_attn_autograd(lstm_model, xb, seed=None if seed is None else seed + i)
            pred = logits.argmax(dim=1)
            logits.gather(1, pred[:, None]).sum().backward()
        out[i:i + len(xb)] = xb.grad.cpu().numpy()
    return out


def compute_channel_importance(lstm_model, X_test, n_samples=100, batch_size=32, channel_names=None):
    """07:203-284 -> DataFrame[Channel, Importance] sorted by importance (normalised to sum 1)."""
    import pandas as pd
    was_training = lstm_model.training
    lstm_model.train()                                   # the reference switches to train mode (07:217-219)
    n_channels = X_test.shape[2]
    if channel_names is None or len(channel_names) != n_channels:
        channel_names = [f"Ch{i + 1}" for i in range(n_channels)]
    n_samples = min(n_samples, len(X_test))
    indices = np.random.choice(len(X_test), n_samples, replace=False)     # same call as 07:233
    grads = input_gradients(lstm_model, X_test[indices], batch_size)
    importance = np.abs(grads).mean(axis=1).sum(axis=0) / n_samples      # |grad| averaged over time, summed over samples
    importance = importance / importance.sum()
    if not was_training:
        lstm_model.eval()
    return pd.DataFrame({"Channel": channel_names, "Importance": importance}).sort_values("Importance", ascending=False)


@torch.no_grad()
def permuted_channel_accuracy(lstm_model, X_subset, y_subset, channels, perms, rows_per_pass=None):
    """Accuracy of the model on V channel-permuted copies of one resident subset (the inner loop of 07:330-345 for all
    (channel, repetition) pairs at once).

    X_subset (n,T,C) float32 numpy / tensor, y_subset (n,) labels, channels (V,) permuted channel per variant (< 0: none),
    perms (V,n) sample order per variant.  The subset is uploaded ONCE; the variants are gathered on the device
    (`bci_permute_channels`) directly in the forward's input layout -- bf16 for the bf16 engine, which rounds x on load anyway --
    and run through the inference forward in full passes of the recurrence kernel instead of in batches of 128.
    Returns (V,) int64 numpy: correctly classified samples per variant (accuracy = count / n, 07:343)."""
    dev = next(lstm_model.parameters()).device
    if dev.type != "cuda":
        raise N.BciError(-1, "permutation importance runs on a CUDA model only (there is no CPU fallback)")
    x = torch.as_tensor(X_subset, dtype=torch.float32).to(dev).contiguous()
    n, T, Cc = (int(v) for v in x.shape)
    channels = np.asarray(channels, dtype=np.int32).reshape(-1)
    V = len(channels)
    perms = np.ascontiguousarray(np.asarray(perms, dtype=np.int32).reshape(V, n))
    if V and (perms.min() < 0 or perms.max() >= n or channels.max() >= Cc):
        raise N.BciError(-1, "permuted_channel_accuracy: perms must index the %d samples and channels the %d channels" % (n, Cc))
    y = torch.as_tensor(np.asarray(y_subset), dtype=torch.int64).to(dev)
    perm_d, ch_d = torch.from_numpy(perms).to(dev).view(-1), torch.from_numpy(channels).to(dev)
    was_training = lstm_model.training
    lstm_model.eval()                                                    # 07:293
    try:
        with torch.cuda.device(dev):
            prec = lstm_model._precision_now()
            hid = lstm_model._engine(prec)
            if rows_per_pass is None:
                rows_per_pass = 2 * ops.lstm_chunk_windows(hid)         # two waves of the recurrence kernel per gather
            total = V * n
            correct = torch.empty((total,), device=dev, dtype=torch.bool)
            for row0 in range(0, total, rows_per_pass):
                rows = min(rows_per_pass, total - row0)
                xv = ops.permute_channels(x, perm_d, ch_d, row0, rows, bf16_out=(prec == "bf16"))
                if prec == "bf16":
                    logits, _, _ = ops.lstm_attn_forward_view(xv.view(-1), hid, rows, T, 0, T * Cc, 0, 0, False)
                else:
                    logits, _, _ = ops.lstm_attn_forward(xv, hid, False)
                idx = torch.arange(row0, row0 + rows, device=dev) % n
                correct[row0:row0 + rows] = logits.argmax(dim=1) == y[idx]          # 07:319,343
                del xv
            counts = correct.view(V, n).sum(dim=1).cpu().numpy()
    finally:
        if was_training:
            lstm_model.train()
    return counts.astype(np.int64)


def compute_permutation_importance(lstm_model, X_test, y_test, n_permutations=5, n_samples=1000, batch_size=128,
                                   channel_names=None):
    """07_explainability.py:287-361 -> DataFrame[Channel, Importance] sorted by importance (mean drop in accuracy when a
    channel's values are shuffled across samples).  Same arguments and the same numpy random calls in the same order (the
    subset draw of 07:304, then one np.random.permutation per (channel, repetition), 07:337) -- the forward passes between them
    draw nothing in eval mode, so seeding numpy as the reference does reproduces its permutations exactly.  `batch_size` is
    accepted for signature compatibility; the 1 + C * n_permutations sweeps run as a few full passes of the B200 forward."""
    import pandas as pd
    n_channels = X_test.shape[2]
    if channel_names is None or len(channel_names) != n_channels:
        channel_names = [f"Ch{i + 1}" for i in range(n_channels)]
    if len(X_test) > n_samples:
        indices = np.random.choice(len(X_test), n_samples, replace=False)
        X_subset, y_subset = X_test[indices], y_test[indices]
    else:
        X_subset, y_subset = X_test, y_test
    n = len(X_subset)
    channels = [-1] + [ch for ch in range(n_channels) for _ in range(n_permutations)]
    perms = [np.arange(n)] + [np.random.permutation(n) for _ in range(n_channels * n_permutations)]
    counts = permuted_channel_accuracy(lstm_model, X_subset, y_subset, channels, np.stack(perms))
    baseline_acc = np.float64(counts[0]) / n                           # == np.mean(pred == y_subset)
    importance_scores = []
    for ch in range(n_channels):
        acc_drops = [baseline_acc - np.float64(c) / n for c in counts[1 + ch * n_permutations:1 + (ch + 1) * n_permutations]]
        importance_scores.append(np.mean(acc_drops))
    return pd.DataFrame({"Channel": channel_names, "Importance": importance_scores}).sort_values("Importance", ascending=False)
