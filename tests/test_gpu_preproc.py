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


# the last two cases span several 8192-sample chunks of the time-parallel filter (warm-up, chunk joins, a partial last tile and chunk)
@pytest.mark.parametrize("R,C,n,dtype", [(3, 61, 5000, np.float64), (2, 7, 1301, np.float32), (1, 64, 256, np.float64),
                                         (2, 5, 20011, np.float32), (1, 35, 41000, np.float64)])
def test_batch_matches_oracle(R, C, n, dtype):
    """Against the oracle (scipy filtfilt + the reference's normalisation / windowing).  Rows that fit one chunk of the time-parallel
    filter start from scipy's own initial state and repeat its operation sequence: <= 1e-9 of the signal's scale.  Longer rows are
    cut into chunks that warm up from a zero state; the warm-up error itself is e^-39, but this 8th-order direct-form band-pass
    (poles at |z| = 0.9954) amplifies ROUNDING differences to ~2e-7 of the output scale -- scipy's own output moves by that much when
    its initial state is perturbed by 1e-15 (oracle check below) -- so no chunked evaluation can agree with it more closely: the
    bound there is 5e-7 of the scale, i.e. the recursion's rounding-noise floor, and 3e-6 on the z-scored fp32 windows."""
    raw = synth.make_raw_eeg(11, R, C, n, dtype=dtype)
    b, a, zi, padlen = pp.design_bandpass()
    out = pp.preprocess_recordings(torch.from_numpy(raw).cuda(), b, a, zi, padlen, want_filtered=True)
    n_seq = (n - 256) // 128 + 1
    assert out["n_seq"] == n_seq and tuple(out["X"].shape) == (R * n_seq, 256, C)
    X = out["X"].cpu().numpy()
    multi = n + 2 * padlen > 8192
    tol_f, tol_s, tol_x = (5e-7, 1e-6, 3e-6) if multi else (1e-9, 1e-9, 2e-6)
    worst = [0.0, 0.0, 0.0]
    for r in range(R):
        filt = po.bandpass_filter(raw[r].astype(np.float64), 1.0, 45.0, 500, 4)
        Xr, _, prm = po.preprocess_recording(raw[r].astype(np.float64), 0)
        worst[0] = max(worst[0], np.abs(out["filtered"][r].cpu().numpy() - filt).max() / np.abs(filt).max())
        worst[1] = max(worst[1], np.abs(out["std"][r].cpu().numpy() / np.asarray(prm["std"]) - 1).max())
        worst[2] = max(worst[2], np.abs(X[r * n_seq:(r + 1) * n_seq] - Xr.astype(np.float32)).max())
    print("preprocess R=%d C=%d n=%d: filtered %.2e of scale, std %.2e, windows %.2e" % (R, C, n, *worst))
    assert worst[0] <= tol_f and worst[1] <= tol_s and worst[2] <= tol_x, worst
    if multi:   # the noise floor claimed above, measured on the oracle itself: scipy's recursion under a 1e-15 initial-state perturbation
        from scipy.signal import lfilter
        x0 = raw[0, 0].astype(np.float64)
        y0, _ = lfilter(b, a, x0, zi=np.asarray(zi) * x0[0])
        y1, _ = lfilter(b, a, x0, zi=np.asarray(zi) * x0[0] * (1 + 1e-15))
        floor = np.abs(y1 - y0).max() / np.abs(y0).max()
        print("  scipy lfilter under a 1e-15 state perturbation moves by %.2e of its scale" % floor)
        assert floor >= 1e-8 and worst[0] <= 10 * floor


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
