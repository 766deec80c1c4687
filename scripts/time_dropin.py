"""Drop-in call contract timing: LSTMODEIntegration.predict_batch(X_numpy, forecast_steps=20, batch_size=512) exactly as
06_lstm_ode_integration.py:801-806 calls it (host numpy in, host numpy out), bf16 mode, N windows from argv."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lstm_ode_bci_b200 import integration, lstm, ode, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 33792
m = lstm.from_params(synth.make_lstm_params(42, 61, 128, 3, logit_gain=12.0), precision=sys.argv[2] if len(sys.argv) > 2 else "bf16")
integ = integration.LSTMODEIntegration(m, ode.CognitiveStateODE(), coupling_strength=0.5)
X = np.random.default_rng(0).standard_normal((N, 256, 61), dtype=np.float32)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    traj, probs, preds = integ.predict_batch(X, forecast_steps=20, batch_size=512, show_progress=False)
    dt = time.perf_counter() - t0
print("predict_batch(%d windows, batch_size=512): %.3f s = %.0f windows/s (pageable numpy in, numpy out); traj %s" % (N, dt, N / dt, traj.shape))
