"""Golden result of the reference's own seeded parameter fit: CognitiveStateODE.fit_to_data (05_ode_model.py:244-322,
differential_evolution(seed=42, maxiter=1000, tol=1e-7, polish=True)), run by the live reference in the build container.

The observations are the exact solution of known rates on t = 0..40 (41 points) from y0 = [0.7, 0.2, 0.1] -- the same
problem tests/test_gpu_next_rows.py fits on the GPU.  Stored: the fitted rates and the final loss.

Run:  python tests/golden/make_golden_fit.py      (about a minute: ~10^4 odeint calls)
"""
import contextlib
import io
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ode_oracle, ref_loader  # noqa: E402

warnings.filterwarnings("ignore")

TRUE_RATES = {"k_ap": 0.12, "k_af": 0.03, "k_pa": 0.2, "k_pf": 0.05, "k_fa": 0.08, "k_fp": 0.15}
Y0 = [0.7, 0.2, 0.1]

if __name__ == "__main__":
    ref05 = ref_loader.load("ref05")
    tp = np.linspace(0, 40, 41)
    k = ode_oracle.rates_to_array(TRUE_RATES)[:, None]
    obs = ode_oracle.exact_solution(ode_oracle.STYLE_REF06, [Y0], k, 40.0, 41)[0]
    ode = ref05.CognitiveStateODE()
    t0 = time.time()
    with contextlib.redirect_stdout(io.StringIO()):
        fitted, loss = ode.fit_to_data(obs, tp)
    print("reference fit: %.1f s, loss %.10f" % (time.time() - t0, loss), fitted)
    order = ["k_ap", "k_af", "k_pa", "k_pf", "k_fa", "k_fp"]
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ode_ref05_fit.npz"),
                        observed=obs, time_points=tp, true_rates=np.array([TRUE_RATES[n] for n in order]),
                        fitted=np.array([fitted[n] for n in order], dtype=np.float64), loss=np.float64(loss), y0=np.array(Y0))
