"""One ODE-ensemble launch pair (for ncu): 16 M coupled trajectories, RK4 S=8, full trajectory."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lstm_ode_bci_b200 import ode, synth
n = 1 << 24
sw = synth.make_ode_sweep(1, n)
dev = {k: torch.tensor(v).cuda() for k, v in sw.items()}
for _ in range(2):
    t, f, _ = ode.solve_ensemble(n, p_open=dev["p_open"], p_closed=dev["p_closed"], rates=dev["rates"], alpha_arr=dev["alpha"],
                                 y0_mode="probs06", coupling=True, substeps=8, want_traj=True)
torch.cuda.synchronize()
print("ok", float(f.sum()))
