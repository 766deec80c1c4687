import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lstm_ode_bci_b200 import ode, ops, synth
def timeit(fn, reps=7, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
peak = ops.fp32_peak_probe(); print("fp32 peak", peak)
n = 1 << 24
sw = synth.make_ode_sweep(1, n)
dev = {k: torch.tensor(v).cuda() for k, v in sw.items()}
for S in (4, 8, 16):
    for wt in (True, False):
        f = lambda: ode.solve_ensemble(n, p_open=dev["p_open"], p_closed=dev["p_closed"], rates=dev["rates"], alpha_arr=dev["alpha"], y0_mode="probs06", coupling=True, substeps=S, want_traj=wt)
        best = timeit(f); flop = 12 + 19 * S * 123 + 220
        print(f"rk4 S={S} traj={wt}: {best:.3f} ms {n/best*1e3/1e9:.2f} Gtraj/s {n/best*1e3*flop/1e12:.1f} TFLOP/s frac {n/best*1e3*flop/1e12/peak:.3f}")
