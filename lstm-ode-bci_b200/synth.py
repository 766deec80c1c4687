"""Deterministic synthetic weights / EEG windows / ODE sweeps (SURVEY.md §8 d).

Everything is drawn from numpy's PCG64 `default_rng(seed)` so that the container that
generated tests/golden/* and the GPU box reproduce identical inputs without shipping
multi-megabyte weight files.  Parameter names and shapes are the reference's state-dict
ABI (04_lstm_model.py:163-204): gate row order i,f,g,o in every (4H, .) matrix.
"""
import math
import numpy as np

DEFAULT_RATES = {  # 05_ode_model.py:87-94
    "k_ap": 0.1, "k_af": 0.02, "k_pa": 0.15, "k_pf": 0.08, "k_fa": 0.05, "k_fp": 0.1,
}
RATE_ORDER = ("k_ap", "k_af", "k_pa", "k_pf", "k_fa", "k_fp")
ALPHA_GRID = (0.0, 0.25, 0.5, 0.75, 1.0)  # 06_lstm_ode_integration.py:531


def lstm_param_shapes(input_size=61, hidden=128, layers=3, classes=2, bidirectional=True, use_attention=True,
                      use_layer_norm=True):
    """Ordered {state_dict key: shape} of EnhancedLSTMModel (04_lstm_model.py:163-204) and, with the switches, of
    AblationLSTMModel (09_sensitivity_analysis.py:176-240: Identity / no attention module / one direction)."""
    H, D = hidden, (2 if bidirectional else 1)
    s = {}
    s["input_proj.0.weight"] = (H, input_size)
    s["input_proj.0.bias"] = (H,)
    if use_layer_norm:
        s["input_proj.1.weight"] = (H,)
        s["input_proj.1.bias"] = (H,)
    for l in range(layers):
        k_in = H if l == 0 else D * H
        for suf in ([""] + (["_reverse"] if bidirectional else [])):
            s[f"lstm.weight_ih_l{l}{suf}"] = (4 * H, k_in)
            s[f"lstm.weight_hh_l{l}{suf}"] = (4 * H, H)
            s[f"lstm.bias_ih_l{l}{suf}"] = (4 * H,)
            s[f"lstm.bias_hh_l{l}{suf}"] = (4 * H,)
    if use_layer_norm:
        s["layer_norm.weight"] = (D * H,)
        s["layer_norm.bias"] = (D * H,)
    if use_attention:
        s["attention.attention.0.weight"] = (D * H // 2, D * H)
        s["attention.attention.0.bias"] = (D * H // 2,)
        s["attention.attention.2.weight"] = (1, D * H // 2)
        s["attention.attention.2.bias"] = (1,)
    s["classifier.0.weight"] = (H, D * H)
    s["classifier.0.bias"] = (H,)
    s["classifier.3.weight"] = (H // 2, H)
    s["classifier.3.bias"] = (H // 2,)
    s["classifier.6.weight"] = (classes, H // 2)
    s["classifier.6.bias"] = (classes,)
    return s


def make_lstm_params(seed=42, input_size=61, hidden=128, layers=3, classes=2,
                     bidirectional=True, logit_gain=1.0, use_attention=True, use_layer_norm=True):
    """fp32 parameters with torch-like init ranges (U(-1/sqrt(fan), 1/sqrt(fan)); LN weight
    ~1, bias ~0 perturbed so the affine terms are exercised).  `logit_gain` scales the last
    classifier layer so P(open)/P(closed) spread past the 0.6 thresholds of
    06_lstm_ode_integration.py:377-382 (default init gives P ~ 0.5, SURVEY.md §8 d)."""
    rng = np.random.default_rng(seed)
    out = {}
    shapes = lstm_param_shapes(input_size, hidden, layers, classes, bidirectional, use_attention, use_layer_norm)
    for name, shape in shapes.items():
        if name in ("input_proj.1.weight", "layer_norm.weight"):
            a = 1.0 + 0.1 * rng.standard_normal(shape)
        elif name in ("input_proj.1.bias", "layer_norm.bias"):
            a = 0.05 * rng.standard_normal(shape)
        else:
            if name.startswith("lstm."):
                bound = 1.0 / math.sqrt(hidden)
            elif name.endswith("weight"):
                bound = 1.0 / math.sqrt(shape[-1])
            else:  # Linear bias: fan_in of the matching weight
                wshape = shapes[name[:-4] + "weight"]
                bound = 1.0 / math.sqrt(wshape[-1])
            a = rng.uniform(-bound, bound, size=shape)
        out[name] = np.ascontiguousarray(a, dtype=np.float32)
    if logit_gain != 1.0:
        out["classifier.6.weight"] = (out["classifier.6.weight"] * np.float32(logit_gain)).astype(np.float32)
    return out


def make_windows(seed, batch, seq_len=256, channels=61, structured=False):
    """x ~ N(0,1) fp32 (B,T,C): reference inputs are per-channel z-scored
    (02_preprocessing.py:134-152).  structured=True adds 1-45 Hz sinusoids at 500 Hz."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((batch, seq_len, channels), dtype=np.float32)
    if structured:
        t = np.arange(seq_len, dtype=np.float32) / 500.0
        f = rng.uniform(1.0, 45.0, size=(batch, 1, channels)).astype(np.float32)
        ph = rng.uniform(0, 2 * np.pi, size=(batch, 1, channels)).astype(np.float32)
        x = (0.6 * x + np.sin(2 * np.pi * f * t[None, :, None] + ph)).astype(np.float32)
    return np.ascontiguousarray(x)


def make_ode_sweep(seed, n):
    """Config 4 (SURVEY.md §8 d): per-trajectory k_af ~ U[0.001,0.2], k_fa ~ U[0.01,0.3]
    (bounds 05_ode_model.py:289,292), other rates default, alpha cycled over the 06:531 grid,
    p_closed ~ U(0,1), p_open = 1 - p_closed.  Returns fp32 SoA arrays."""
    rng = np.random.default_rng(seed)
    rates = np.empty((6, n), dtype=np.float32)
    for i, k in enumerate(RATE_ORDER):
        rates[i, :] = DEFAULT_RATES[k]
    rates[1, :] = rng.uniform(0.001, 0.2, size=n)
    rates[4, :] = rng.uniform(0.01, 0.3, size=n)
    alpha = np.asarray(ALPHA_GRID, dtype=np.float32)[np.arange(n) % len(ALPHA_GRID)]
    p_closed = rng.uniform(0.0, 1.0, size=n).astype(np.float32)
    p_open = (np.float32(1.0) - p_closed).astype(np.float32)
    return {"rates": rates, "alpha": np.ascontiguousarray(alpha), "p_open": p_open, "p_closed": p_closed}


def make_raw_eeg(seed, recordings, channels=61, samples=150000, fs=500.0, dtype=np.float64):
    """Raw-recording stand-in for mne's raw.get_data() (02_preprocessing.py:200): volts-scale (R, C, n) with a slow drift
    (below the 1 Hz corner), 50 Hz mains (above the 45 Hz corner), a 10 Hz alpha rhythm and broadband noise, a different
    DC offset and gain per channel, so the band-pass, the z-score and the windowing all have something to do."""
    rng = np.random.default_rng(seed)
    t = np.arange(samples, dtype=np.float64) / fs
    x = 8e-6 * rng.standard_normal((recordings, channels, samples))
    gain = rng.uniform(0.5, 2.0, size=(recordings, channels, 1))
    dc = 1e-4 * rng.standard_normal((recordings, channels, 1))
    ph = rng.uniform(0, 2 * np.pi, size=(recordings, channels, 3, 1))
    x += 2e-5 * np.sin(2 * np.pi * 0.2 * t + ph[:, :, 0])
    x += 1e-5 * np.sin(2 * np.pi * 50.0 * t + ph[:, :, 1])
    x += 1.2e-5 * np.sin(2 * np.pi * 10.0 * t + ph[:, :, 2])
    return np.ascontiguousarray((gain * x + dc).astype(dtype))
