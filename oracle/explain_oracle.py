"""CPU restatement of the reference's permutation importance and ODE sensitivity analysis (TEST INFRASTRUCTURE ONLY -- imported
by tests/ only, never by the product package).

`permutation_importance` follows 07_explainability.py:287-361 step by step (subset draw, batched argmax predictions, one
np.random.permutation per (channel, repetition), accuracy drop averaged over repetitions) around a caller-supplied
`predict(X) -> class ids`; `permuted_copy` is the host-side gather of 07:336-339 that the CUDA kernel bci_permute_channels
replaces.  `sensitivity` follows 05_ode_model.py:687-719 around a caller-supplied steady-state function.
Pinned against the live reference by tests/golden/explain_ref07.npz and ode_ref05_sensitivity.npz
(made by tests/golden/make_golden_explain.py).
"""
import numpy as np


def permuted_copy(X, perm_idx, ch_idx):
    """07:336-339."""
    Xp = X.copy()
    Xp[:, :, ch_idx] = X[perm_idx, :, ch_idx]
    return Xp


def permutation_importance(predict, X_test, y_test, n_permutations=5, n_samples=1000):
    """-> (importance per channel in channel order, baseline accuracy).  Draws from numpy's global generator in the
    reference's order, so np.random.seed(s) before the call reproduces the reference run seeded the same way."""
    n_channels = X_test.shape[2]
    if len(X_test) > n_samples:                                      # 07:303-309
        indices = np.random.choice(len(X_test), n_samples, replace=False)
        X_subset, y_subset = X_test[indices], y_test[indices]
    else:
        X_subset, y_subset = X_test, y_test
    baseline_acc = np.mean(predict(X_subset) == y_subset)            # 07:325-326
    scores = []
    for ch_idx in range(n_channels):                                 # 07:334-349
        drops = []
        for _ in range(n_permutations):
            perm_idx = np.random.permutation(len(X_subset))
            drops.append(baseline_acc - np.mean(predict(permuted_copy(X_subset, perm_idx, ch_idx)) == y_subset))
        scores.append(np.mean(drops))
    return np.array(scores, dtype=np.float64), float(baseline_acc)


def sensitivity(base_params, steady_state, perturbation=0.2):
    """05:693-713: central difference of the steady state over +-20 % of each rate -> (n_params, 3)."""
    rows = []
    for name in base_params:
        pair = []
        for factor in (1 - perturbation, 1 + perturbation):
            p = dict(base_params)
            p[name] = base_params[name] * factor
            pair.append(np.asarray(steady_state(p), dtype=np.float64))
        rows.append((pair[1] - pair[0]) / (2 * perturbation * base_params[name]))
    return np.array(rows)
