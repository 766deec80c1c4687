"""SURVEY.md §8 f rank 1: gradient-based attribution on the B200 backward pass.

`compute_channel_importance` mirrors 07_explainability.py:203-284 (same arguments, same sampling call, same
DataFrame).  The reference back-propagates outputs[i, pred_i] once PER SAMPLE with retain_graph=True -- B full backward
passes per batch.  Windows are independent, so d outputs[i, pred_i] / d x[i] for every i of a batch is exactly row i of ONE
backward pass with dlogits[i, pred_i] = 1: `input_gradients` does that (B x fewer BPTT launches).  The per-sample pattern
of the reference also works unchanged against lstm.EnhancedLSTMModel (autograd bridge keeps the saved forward)."""
import numpy as np
import torch

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
