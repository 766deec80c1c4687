"""Golden vectors for CognitiveStateODE.solve_with_modulation (05_ode_model.py:171-196), produced by the live reference
(build container only; needs /root/reference).  The modulation functions are oracle.ode_oracle.modulation_cases().

Run:  python tests/golden/make_golden_modulation.py
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ode_oracle, ref_loader  # noqa: E402

warnings.filterwarnings("ignore")

if __name__ == "__main__":
    ref05 = ref_loader.load("ref05")
    out = {}
    for name, fn, y0, t_span, n_points in ode_oracle.modulation_cases():
        ode = ref05.CognitiveStateODE()          # default rates 05:87-94
        t, sol = ode.solve_with_modulation(y0, t_span, fn, n_points=n_points)
        out[name + "_t"], out[name + "_sol"] = t, sol
        print(name, sol.shape, sol[-1])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ode_ref05_modulation.npz"), **out)
