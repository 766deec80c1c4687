"""Golden vectors for the SURVEY.md §8 f rows (build container only; needs /root/reference):

  preproc_ref02.npz   02_preprocessing.py's own bandpass_filter / normalize_data / create_sequences on a seeded raw recording
  ablation_*.npz      09_sensitivity_analysis.py's AblationLSTMModel (forward logits + autograd gradient summaries)

Run:  python tests/golden/make_golden_next.py     (inputs are regenerated from seeds by lstm_ode_bci_b200.synth)
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from lstm_ode_bci_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
warnings.filterwarnings("ignore")


def preproc_case():
    ref02 = ref_loader.load("ref02")
    seed, C, n = 7, 61, 3000
    raw = synth.make_raw_eeg(seed, 1, C, n)[0]
    filt = ref02.bandpass_filter(raw, ref02.LOWCUT, ref02.HIGHCUT, ref02.SAMPLING_RATE, ref02.FILTER_ORDER)
    norm, mean, std = ref02.normalize_data(filt.copy())
    X, y = ref02.create_sequences(norm, 1, ref02.SEQUENCE_LENGTH, ref02.SEQUENCE_OVERLAP)
    # second recording normalised with the first one's statistics (02:207-210)
    raw2 = synth.make_raw_eeg(seed + 1, 1, C, n)[0]
    filt2 = ref02.bandpass_filter(raw2, ref02.LOWCUT, ref02.HIGHCUT, ref02.SAMPLING_RATE, ref02.FILTER_ORDER)
    norm2, _, _ = ref02.normalize_data(filt2.copy(), mean.reshape(-1, 1), std.reshape(-1, 1))
    X2, _ = ref02.create_sequences(norm2, 0, ref02.SEQUENCE_LENGTH, ref02.SEQUENCE_OVERLAP)
    np.savez_compressed(os.path.join(OUT, "preproc_ref02.npz"), seed=seed, C=C, n=n,
                        filtered=filt[::7, :], mean=mean, std=std, X=X.astype(np.float32)[:, :, ::5], y=y,
                        X2=X2.astype(np.float32)[::3, :, ::9], n_seq=len(X))
    print("preproc: filtered max", np.abs(filt).max(), "X", X.shape, "std[0]", std[0])


if __name__ == "__main__":
    which = sys.argv[1:] or ["preproc", "ablation"]
    if "preproc" in which:
        preproc_case()
    if "ablation" in which:
        from make_golden_ablation import ablation_cases  # noqa: E402
        ablation_cases()
