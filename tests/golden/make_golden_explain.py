"""Golden vectors for the callers next to the hot path that SURVEY.md §8 f ranks 1 and 2 name beside the rows already covered
(build container only; needs /root/reference):

  explain_ref07.npz        07_explainability.py's own compute_permutation_importance (07:287-361) on a seeded model / test set
  ode_ref05_sensitivity.npz  05_ode_model.py's own sensitivity_analysis (05:687-750) and get_steady_state (05:198-221)

Run:  python tests/golden/make_golden_explain.py     (inputs are regenerated from seeds by lstm_ode_bci_b200.synth; ~1 minute)
"""
import contextlib
import io
import os
import sys
import warnings
from unittest import mock

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from lstm_ode_bci_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
warnings.filterwarnings("ignore")

# the permutation-importance case (tests/test_gpu_next_rows.py regenerates the same inputs)
PERM_CASE = dict(param_seed=31, logit_gain=200.0, x_seed=32, n_test=96, n_samples=64, n_permutations=2, numpy_seed=7,
                 label_flip_every=5, boosted_channels=[3, 17, 40], boost=30.0)


def perm_case_params(final_bias=None):
    """Random-init weights barely look at any single channel and put every window in one class (SURVEY.md §8 d: P ~ 0.5): three
    input-projection columns are amplified so those channels matter, and the last bias is shifted so the decision boundary runs
    through the middle of the test set's logit margins (the shift is stored in the fixture as `final_bias`)."""
    c = PERM_CASE
    params = synth.make_lstm_params(c["param_seed"], 61, 128, 3, logit_gain=c["logit_gain"])
    params["input_proj.0.weight"][:, c["boosted_channels"]] *= np.float32(c["boost"])
    if final_bias is not None:
        params["classifier.6.bias"] = np.asarray(final_bias, dtype=np.float32)
    return params


def permutation_case():
    ref07 = ref_loader.load("ref07")
    c = PERM_CASE
    params = perm_case_params()
    model = ref07.EnhancedLSTMModel(61, 128, 3, 2, 0.4, True)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    model.eval()
    X = synth.make_windows(c["x_seed"], c["n_test"], 256, 61, structured=True)
    with torch.no_grad():
        logits = model(torch.from_numpy(X)).numpy()
    final_bias = params["classifier.6.bias"].copy()
    m = np.sort(logits[:, 1] - logits[:, 0])
    mid = m[len(m) // 2 - 5:len(m) // 2 + 5]            # boundary in the widest gap between the ten central margins: no
    j = int(np.argmax(np.diff(mid)))                    # unpermuted window sits on the decision boundary
    final_bias[1] -= np.float32(0.5 * (mid[j] + mid[j + 1]))
    params = perm_case_params(final_bias)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    with torch.no_grad():
        logits = model(torch.from_numpy(X)).numpy()
    y = logits.argmax(1)
    y[::c["label_flip_every"]] ^= 1                    # an imperfect classifier: baseline accuracy 0.8, room to move both ways
    np.random.seed(c["numpy_seed"])
    with contextlib.redirect_stdout(io.StringIO()):
        df = ref07.compute_permutation_importance(model, X, y, n_permutations=c["n_permutations"], n_samples=c["n_samples"],
                                                  batch_size=32)
    names = list(ref07.EEG_CHANNELS)
    order = np.array([names.index(ch) for ch in df["Channel"]], dtype=np.int32)        # the DataFrame's (sorted) row order
    imp = np.empty(61, dtype=np.float64)
    imp[order] = df["Importance"].to_numpy()
    margin = np.abs(logits[:, 1] - logits[:, 0])
    np.savez_compressed(os.path.join(OUT, "explain_ref07.npz"), labels=y.astype(np.int64), importance=imp, sorted_order=order,
                        logits=logits, final_bias=final_bias, **{k: np.asarray(v) for k, v in c.items()})
    print("permutation importance: max %.4f min %.4f nonzero %d, smallest logit margin %.3g"
          % (imp.max(), imp.min(), int((imp != 0).sum()), margin.min()))


SENS_PARAMS = [None,  # the reference's defaults (05:87-94)
               {"k_ap": 0.12, "k_af": 0.03, "k_pa": 0.2, "k_pf": 0.05, "k_fa": 0.08, "k_fp": 0.15}]


def sensitivity_case():
    ref05 = ref_loader.load("ref05")
    out = {}
    fake_plt = mock.MagicMock()
    fake_plt.subplots.return_value = (mock.MagicMock(), mock.MagicMock())
    for j, p in enumerate(SENS_PARAMS):
        ode = ref05.CognitiveStateODE(None if p is None else dict(p))
        with mock.patch.object(ref05, "plt", fake_plt), contextlib.redirect_stdout(io.StringIO()):
            res = ref05.sensitivity_analysis(ode, mock.MagicMock())      # the figure goes to the mocked pyplot
        out["names_%d" % j] = np.array([r["parameter"] for r in res])
        out["sens_%d" % j] = np.array([[r["sens_Active"], r["sens_Passive"], r["sens_Fatigued"]] for r in res], dtype=np.float64)
        ss = ode.get_steady_state()
        out["steady_%d" % j] = np.array([ss["Active"], ss["Passive"], ss["Fatigued"]], dtype=np.float64)
        out["params_%d" % j] = np.array([ode.params[k] for k in synth.RATE_ORDER], dtype=np.float64)
        print("sensitivity case %d: |sens| max %.4f, steady %s" % (j, np.abs(out["sens_%d" % j]).max(), out["steady_%d" % j]))
    np.savez_compressed(os.path.join(OUT, "ode_ref05_sensitivity.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["permutation", "sensitivity"]
    if "sensitivity" in which:
        sensitivity_case()
    if "permutation" in which:
        permutation_case()
