"""Host-side mirror of the reference's preprocessing functions (02_preprocessing.py:114-221) over the CUDA path.

`bandpass_filter`, `normalize_data`, `create_sequences` keep the reference's names, arguments and return values (numpy in,
numpy out -- or CUDA tensors in, CUDA tensors out), and `preprocess_recordings` is the fused call the three collapse
into on the GPU: raw (R, C, n) recordings -> (R * n_seq, 256, 61) fp32 windows ready for EnhancedLSTMModel, one pass,
every filtered sample written straight into the (up to two) windows that contain it.

Filter design (scipy.signal.butter / lfilter_zi: nine coefficients, computed once on the host exactly as the reference
does at 02:128-130) is not on the hot path and stays with scipy; the recursion, statistics and windowing run in
`bci_preprocess` (csrc/preproc.cu).  There is no CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _native as N
from .ops import _ptr, _stream

SAMPLING_RATE, SEQUENCE_LENGTH, SEQUENCE_OVERLAP = 500, 256, 0.5     # 02:49-51
LOWCUT, HIGHCUT, FILTER_ORDER = 1.0, 45.0, 4                          # 02:52-54


def design_bandpass(lowcut=LOWCUT, highcut=HIGHCUT, fs=SAMPLING_RATE, order=FILTER_ORDER):
    """(b, a, zi, padlen) exactly as bandpass_filter (02:126-131) and scipy.signal.filtfilt's defaults derive them."""
    from scipy.signal import butter, lfilter_zi
    nyq = 0.5 * fs
    b, a = butter(order, [lowcut / nyq, highcut / nyq], btype="band")
    zi = lfilter_zi(b, a)
    return np.ascontiguousarray(b, np.float64), np.ascontiguousarray(a, np.float64), np.ascontiguousarray(zi, np.float64), \
        3 * max(len(a), len(b))


def _as_cuda(data):
    if isinstance(data, torch.Tensor):
        if not data.is_cuda:
            raise N.BciError(-1, "preprocessing runs on CUDA tensors (or numpy arrays, which are copied to the GPU); no CPU fallback")
        return data, True
    return torch.from_numpy(np.ascontiguousarray(data)).cuda(), False


def preprocess_recordings(raw, b, a, zi=None, padlen=None, seq_len=SEQUENCE_LENGTH, overlap=SEQUENCE_OVERLAP, mean=None,
                          std=None, want_filtered=False):
    """raw (R, C, n) or (C, n), fp64 or fp32 (numpy or CUDA tensor) ->
    dict(X (R*n_seq, seq_len, C) fp32 CUDA, mean (R, C), std (R, C) fp64 CUDA, filtered (R, C, n) fp64 | None, n_seq)."""
    x, _ = _as_cuda(raw)
    if x.dim() == 2:
        x = x[None]
    if x.dtype not in (torch.float32, torch.float64):
        x = x.double()
    x = x.contiguous()
    R, Cc, n = (int(v) for v in x.shape)
    b = np.ascontiguousarray(b, np.float64)
    a = np.ascontiguousarray(a, np.float64)
    if len(a) != len(b):
        m = max(len(a), len(b))
        b, a = np.pad(b, (0, m - len(b))), np.pad(a, (0, m - len(a)))
    if zi is None:
        from scipy.signal import lfilter_zi
        zi = lfilter_zi(b, a)
    zi = np.ascontiguousarray(zi, np.float64)
    if padlen is None:
        padlen = 3 * len(a)
    if n <= padlen:
        raise ValueError("The length of the input vector x must be greater than padlen, which is %d." % padlen)  # scipy's error
    step = int(seq_len * (1 - overlap))
    args = N.PreprocArgs()
    args.n_recordings, args.n_channels, args.n_samples = R, Cc, n
    args.in_dtype = N.OUT_F64 if x.dtype == torch.float64 else N.OUT_F32
    args.order = len(a) - 1
    args.b_host = b.ctypes.data_as(C.POINTER(C.c_double))
    args.a_host = a.ctypes.data_as(C.POINTER(C.c_double))
    args.zi_host = zi.ctypes.data_as(C.POINTER(C.c_double))
    args.padlen, args.seq_len, args.step = int(padlen), int(seq_len), step
    keep = []
    if mean is not None:
        mt = torch.as_tensor(np.asarray(mean, np.float64).reshape(-1)).cuda() if not isinstance(mean, torch.Tensor) else mean.double().reshape(-1).contiguous()
        st = torch.as_tensor(np.asarray(std, np.float64).reshape(-1)).cuda() if not isinstance(std, torch.Tensor) else std.double().reshape(-1).contiguous()
        if mt.numel() != Cc or st.numel() != Cc:
            raise N.BciError(-1, "mean/std must have one entry per channel")
        keep += [mt, st]
        args.mean_in, args.std_in = mt.data_ptr(), st.data_ptr()
    n_seq = (n - seq_len) // step + 1 if n >= seq_len else 0
    if n_seq <= 0:
        raise N.BciError(-1, "recording shorter than one window")
    X = torch.empty((R * n_seq, seq_len, Cc), device=x.device, dtype=torch.float32)
    mo = torch.empty((R, Cc), device=x.device, dtype=torch.float64)
    so = torch.empty((R, Cc), device=x.device, dtype=torch.float64)
    filt = torch.empty((R, Cc, n), device=x.device, dtype=torch.float64) if want_filtered else None
    nb = C.c_size_t(0)
    N.check(N.lib().bci_preprocess_workspace_bytes(C.byref(args), C.byref(nb)))
    ws = torch.empty((nb.value,), device=x.device, dtype=torch.uint8)
    with torch.cuda.device(x.device):
        N.check(N.lib().bci_preprocess(C.byref(args), _ptr(x), _ptr(X), _ptr(mo), _ptr(so), _ptr(filt), _ptr(ws), nb.value, _stream()))
    return {"X": X, "mean": mo, "std": so, "filtered": filt, "n_seq": n_seq}


# ---- the reference's three functions, same signatures (02:114-180) -----------------------------------------------------
def bandpass_filter(data, lowcut, highcut, fs, order=4):
    """data (n_channels, n_samples) -> filtered, same shape, float64 (02:114-131)."""
    b, a, zi, padlen = design_bandpass(lowcut, highcut, fs, order)
    x, was_tensor = _as_cuda(data)
    n = int(x.shape[-1])
    out = preprocess_recordings(x, b, a, zi, padlen, seq_len=min(256, n), overlap=0.0, want_filtered=True)["filtered"][0]
    return out if was_tensor else out.cpu().numpy()


def normalize_data(data, mean=None, std=None):
    """Per-channel z-score (02:134-154): returns (normalized, mean (C,), std (C,))."""
    x, was_tensor = _as_cuda(data)
    x = x.double()
    mu = x.mean(dim=1, keepdim=True) if mean is None else torch.as_tensor(np.asarray(mean, np.float64)).reshape(-1, 1).to(x.device)
    if std is None:
        sd = x.std(dim=1, keepdim=True, unbiased=False)
        sd = torch.where(sd < 1e-10, torch.full_like(sd, 1e-10), sd)
    else:
        sd = torch.as_tensor(np.asarray(std, np.float64)).reshape(-1, 1).to(x.device)
    out = (x - mu) / sd
    if was_tensor:
        return out, mu.flatten(), sd.flatten()
    return out.cpu().numpy(), mu.flatten().cpu().numpy(), sd.flatten().cpu().numpy()


def create_sequences(data, label, seq_length, overlap):
    """data (n_channels, n_samples) -> X (n_sequences, seq_length, n_channels), y (n_sequences,) (02:157-180)."""
    x, was_tensor = _as_cuda(data)
    step = int(seq_length * (1 - overlap))
    X = x.unfold(1, seq_length, step).permute(1, 2, 0).contiguous()     # (n_seq, seq_length, C)
    y = torch.full((X.shape[0],), int(label), dtype=torch.int64, device=x.device)
    return (X, y) if was_tensor else (X.cpu().numpy(), y.cpu().numpy())


def preprocess_recording(data, label, normalization_params=None, lowcut=LOWCUT, highcut=HIGHCUT, fs=SAMPLING_RATE,
                         order=FILTER_ORDER, seq_len=SEQUENCE_LENGTH, overlap=SEQUENCE_OVERLAP):
    """load_and_preprocess_recording (02:183-217) after the mne load: data (C, n) -> (X fp32 CUDA, y, normalization_params)."""
    b, a, zi, padlen = design_bandpass(lowcut, highcut, fs, order)
    mean = std = None
    if normalization_params:
        mean, std = normalization_params["mean"], normalization_params["std"]
    out = preprocess_recordings(data, b, a, zi, padlen, seq_len, overlap, mean, std)
    y = torch.full((out["X"].shape[0],), int(label), dtype=torch.int64, device=out["X"].device)
    if not normalization_params:
        normalization_params = {"mean": out["mean"][0].cpu().tolist(), "std": out["std"][0].cpu().tolist()}
    return out["X"], y, normalization_params
