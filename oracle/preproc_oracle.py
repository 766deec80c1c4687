"""CPU restatement of the reference's preprocessing (TEST INFRASTRUCTURE ONLY -- imported by tests/, smoke() and
bench.py's CPU legs, never by the product package).

Follows 02_preprocessing.py:114-180 line by line (bandpass_filter, normalize_data, create_sequences).  The arithmetic
of `filtfilt` lives in scipy (reference requirement `scipy>=1.11.0`, unpinned; 1.18.1 in the build container), so its
published algorithm is restated here twice: `filtfilt_restated` spells out the padding / initial-state / reversal steps
around scipy.signal.lfilter, and `lfilter_df2t` is the direct-form-II-transposed recursion itself in plain numpy
(small cases).  Pinned against the live reference by tests/golden/preproc_ref02.npz (made by tests/golden/make_golden.py).
"""
import numpy as np


def butter_band(lowcut, highcut, fs, order=4):
    """02:126-130."""
    from scipy.signal import butter
    nyq = 0.5 * fs
    return butter(order, [lowcut / nyq, highcut / nyq], btype="band")


def lfilter_df2t(b, a, x, zi):
    """scipy's _linear_filter for 1-D x: y = z0 + b0 x; z_j = z_{j+1} + b_{j+1} x - a_{j+1} y.  Returns (y, z_final)."""
    b = np.asarray(b, np.float64) / a[0]
    a = np.asarray(a, np.float64) / a[0]
    order = len(a) - 1
    z = np.array(zi, np.float64).copy()
    y = np.empty(len(x), np.float64)
    for k, xv in enumerate(np.asarray(x, np.float64)):
        yv = z[0] + b[0] * xv
        for j in range(order - 1):
            z[j] = z[j + 1] + b[j + 1] * xv - a[j + 1] * yv
        z[order - 1] = b[order] * xv - a[order] * yv
        y[k] = yv
    return y, z


def odd_ext(x, n):
    """scipy.signal._arraytools.odd_ext along the last axis."""
    left = 2 * x[..., :1] - x[..., n:0:-1]
    right = 2 * x[..., -1:] - x[..., -2:-(n + 2):-1]
    return np.concatenate([left, x, right], axis=-1)


def filtfilt_restated(b, a, x, use_numpy_recursion=False):
    """scipy.signal.filtfilt(b, a, x, axis=-1) with its defaults (padtype='odd', padlen=3*max(len(a),len(b)), method='pad')."""
    from scipy.signal import lfilter, lfilter_zi
    x = np.asarray(x, np.float64)
    edge = 3 * max(len(a), len(b))
    if x.shape[-1] <= edge:
        raise ValueError("The length of the input vector x must be greater than padlen, which is %d." % edge)
    ext = odd_ext(x, edge)
    zi = lfilter_zi(b, a)
    if use_numpy_recursion:
        rows = ext.reshape(-1, ext.shape[-1])
        out = np.empty_like(rows)
        for r in range(rows.shape[0]):
            y, _ = lfilter_df2t(b, a, rows[r], zi * rows[r, 0])
            y2, _ = lfilter_df2t(b, a, y[::-1], zi * y[-1])
            out[r] = y2[::-1]
        y = out.reshape(ext.shape)
    else:
        shape = [1] * x.ndim
        shape[-1] = zi.size
        ziv = zi.reshape(shape)
        y, _ = lfilter(b, a, ext, axis=-1, zi=ziv * ext[..., :1])
        y, _ = lfilter(b, a, y[..., ::-1], axis=-1, zi=ziv * y[..., -1:])
        y = y[..., ::-1]
    return y[..., edge:-edge]


def bandpass_filter(data, lowcut, highcut, fs, order=4):
    """02:114-131."""
    b, a = butter_band(lowcut, highcut, fs, order)
    return filtfilt_restated(b, a, data)


def normalize_data(data, mean=None, std=None):
    """02:134-154."""
    if mean is None:
        mean = np.mean(data, axis=1, keepdims=True)
    if std is None:
        std = np.std(data, axis=1, keepdims=True)
        std[std < 1e-10] = 1e-10
    return (data - mean) / std, np.asarray(mean).flatten(), np.asarray(std).flatten()


def create_sequences(data, label, seq_length, overlap):
    """02:157-180."""
    n_channels, n_samples = data.shape
    step = int(seq_length * (1 - overlap))
    starts = range(0, n_samples - seq_length + 1, step)
    X = np.array([data[:, s:s + seq_length].T for s in starts])
    return X, np.full(len(X), label)


def preprocess_recording(data, label, normalization_params=None, lowcut=1.0, highcut=45.0, fs=500, order=4, seq_len=256,
                         overlap=0.5):
    """load_and_preprocess_recording (02:183-217) after the mne load."""
    data = bandpass_filter(data, lowcut, highcut, fs, order)
    if normalization_params:
        mean = np.array(normalization_params["mean"]).reshape(-1, 1)
        std = np.array(normalization_params["std"]).reshape(-1, 1)
        data, _, _ = normalize_data(data, mean, std)
    else:
        data, mean, std = normalize_data(data)
        normalization_params = {"mean": mean.tolist(), "std": std.tolist()}
    X, y = create_sequences(data, label, seq_len, overlap)
    return X, y, normalization_params
