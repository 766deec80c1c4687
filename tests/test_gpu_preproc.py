"""GPU parity of the preprocessing path (SURVEY.md §8 f row 4; 02_preprocessing.py:114-221): `bci_preprocess` through
the host mirror against the oracle (oracle/preproc_oracle.py) and the golden outputs of the live reference.
Tolerances: fp64 band-passed signal <= 1e-9 of its max-abs (the recursion is ill-conditioned at the 1 Hz corner: both
sides are fp64 but the GPU contracts to FMA), statistics <= 1e-9 relative, fp32 windows <= 2e-6."""
import numpy as np
import pytest
import torch

from lstm_ode_bci_b200 import preprocessing as pp, synth
from oracle import preproc_oracle as po

pytestmark = pytest.mark.gpu


def test_matches_reference_golden(golden):
    g = golden("preproc_ref02.npz")
    seed, C, n = int(g["seed"]), int(g["C"]), int(g["n"])
    raw = synth.make_raw_eeg(seed, 1, C, n)[0]
    filt = pp.bandpass_filter(raw, 1.0, 45.0, 500, 4)
    assert filt.dtype == np.float64 and filt.shape == raw.shape
    assert np.abs(filt[::7] - g["filtered"]).max() <= 1e-9 * np.abs(g["filtered"]).max()
    X, y, prm = pp.preprocess_recording(raw, 1)
    assert tuple(X.shape) == (int(g["n_seq"]), 256, C) and X.dtype == torch.float32
    assert np.array_equal(y.cpu().numpy(), g["y"])
    assert np.abs(np.asarray(prm["std"]) / g["std"] - 1).max() <= 1e-9
    assert np.abs(np.asarray(prm["mean"]) - g["mean"]).max() <= 1e-9 * g["std"].max()
    assert np.abs(X.cpu().numpy()[:, :, ::5] - g["X"]).max() <= 2e-6
    raw2 = synth.make_raw_eeg(seed + 1, 1, C, n)[0]
    X2, _, _ = pp.preprocess_recording(raw2, 0, prm)
    assert np.abs(X2.cpu().numpy()[::3, :, ::9] - g["X2"]).max() <= 2e-6


@pytest.mark.parametrize("R,C,n,dtype", [(3, 61, 5000, np.float64), (2, 7, 1301, np.float32), (1, 64, 256, np.float64)])
def test_batch_matches_oracle(R, C, n, dtype):
    raw = synth.make_raw_eeg(11, R, C, n, dtype=dtype)
    b, a, zi, padlen = pp.design_bandpass()
    out = pp.preprocess_recordings(torch.from_numpy(raw).cuda(), b, a, zi, padlen, want_filtered=True)
    n_seq = (n - 256) // 128 + 1
    assert out["n_seq"] == n_seq and tuple(out["X"].shape) == (R * n_seq, 256, C)
    X = out["X"].cpu().numpy()
    for r in range(R):
        filt = po.bandpass_filter(raw[r].astype(np.float64), 1.0, 45.0, 500, 4)
        assert np.abs(out["filtered"][r].cpu().numpy() - filt).max() <= 1e-9 * np.abs(filt).max()
        Xr, _, prm = po.preprocess_recording(raw[r].astype(np.float64), 0)
        assert np.abs(out["std"][r].cpu().numpy() / np.asarray(prm["std"]) - 1).max() <= 1e-9
        assert np.abs(X[r * n_seq:(r + 1) * n_seq] - Xr.astype(np.float32)).max() <= 2e-6


def test_reference_function_mirrors_and_window_geometry():
    raw = synth.make_raw_eeg(5, 1, 9, 1000)[0]
    filt = po.bandpass_filter(raw, 1.0, 45.0, 500, 4)
    norm, mean, std = pp.normalize_data(filt)
    nr, mr, sr = po.normalize_data(filt.copy())
    assert np.abs(norm - nr).max() <= 1e-10 and np.abs(mean - mr).max() <= 1e-18 and np.abs(std / sr - 1).max() <= 1e-12
    for L, ov in ((256, 0.5), (100, 0.75), (64, 0.0)):
        X, y = pp.create_sequences(norm, 1, L, ov)
        Xs, ys = po.create_sequences(norm, 1, L, ov)          # same input: pure data movement, bit-exact
        assert X.shape == Xs.shape and np.array_equal(X, Xs) and np.array_equal(y, ys)
        Xr, yr = po.create_sequences(nr, 1, L, ov)
        b, a, zi, padlen = pp.design_bandpass()
        out = pp.preprocess_recordings(raw, b, a, zi, padlen, seq_len=L, overlap=ov)
        assert np.abs(out["X"].cpu().numpy() - Xr.astype(np.float32)).max() <= 2e-6


def test_errors_match_scipy_and_no_cpu_path():
    from lstm_ode_bci_b200._native import BciError
    b, a, zi, padlen = pp.design_bandpass()
    with pytest.raises(ValueError):
        pp.preprocess_recordings(np.zeros((2, 27)), b, a, zi, padlen)
    with pytest.raises(BciError):
        pp.bandpass_filter(torch.zeros(2, 500), 1.0, 45.0, 500)     # CPU tensor: no fallback


def test_windows_feed_the_model():
    """raw recording -> windows -> EnhancedLSTMModel probabilities, all on the device (config 5 fed from raw data)."""
    from lstm_ode_bci_b200 import lstm
    from oracle import lstm_oracle
    raw = synth.make_raw_eeg(21, 1, 61, 256 * 3)[0]
    X, _, _ = pp.preprocess_recording(raw, 0)
    params = synth.make_lstm_params(42, 61, 128, 3, logit_gain=8.0)
    m = lstm.from_params(params, precision="fp32")
    probs = m.predict_proba(X).cpu().numpy()
    Xo, _, _ = po.preprocess_recording(raw, 0)
    want = lstm_oracle.softmax_probs(lstm_oracle.forward(params, Xo.astype(np.float32))[0])
    assert np.abs(probs - want).max() <= 2e-5
