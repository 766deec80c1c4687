"""BASELINE configs[3]: three-state ODE ensemble, N = 1M ... 64M coupled trajectories with swept k_af / k_fa / alpha / P(closed)
(synth.make_ode_sweep), RK4 (S = 8 sub-steps per output interval) and RK45 (rtol 1e-3, atol 1e-6); CUDA-event timing, best of 5."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lstm_ode_bci_b200 import ode, ops, synth
S = 8
FLOP = 12 + 19 * S * 123 + 220
peak = ops.fp32_peak_probe()
out = {"fp32_fma_peak_tflops": peak, "substeps": S, "flop_per_trajectory": FLOP, "rows": []}
def best_ms(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
for logn in (20, 22, 24, 26):
    n = 1 << logn
    sw = synth.make_ode_sweep(42, n)
    dev = {k: torch.from_numpy(v).cuda() for k, v in sw.items()}
    for mode, traj in (("rk4", True), ("rk4", False), ("rk45", False)):
        if traj and n * 240 > 12e9:
            continue
        f = lambda: ode.solve_ensemble(n, p_open=dev["p_open"], p_closed=dev["p_closed"], rates=dev["rates"], alpha_arr=dev["alpha"],
                                       y0_mode="probs06", coupling=True, style="ref06", mode=mode, t_end=20.0, n_points=20,
                                       substeps=S, want_traj=traj)
        ms = best_ms(f)
        row = {"n": n, "mode": mode, "full_trajectory": traj, "ms": ms, "trajectories_per_s": n / ms * 1e3}
        if mode == "rk4":
            row["tflops"] = n * FLOP / ms / 1e9
            row["frac_of_fp32_peak"] = row["tflops"] / peak
        out["rows"].append(row); print(row, flush=True)
    del dev
json.dump(out, open("gpurun_out/sweep_config4.json", "w"), indent=1)
