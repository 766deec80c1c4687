"""Quick stage timings on the GPU (development aid, not the bench)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lstm_ode_bci_b200 import lstm, ode, ops, synth

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))

print("fp32 FMA probe TFLOP/s:", ops.fp32_peak_probe())
for prec in sys.argv[1:] or ["fp32"]:
    for H in (128,):
        params = synth.make_lstm_params(42, 61, H, 3)
        m = lstm.from_params(params, precision=prec)
        for B in (256, 2048) if prec == "fp32" else (4096, 16384):
            x = torch.randn(B, 256, 61, device="cuda")
            best, med = timeit(lambda: m.predict_proba(x), reps=3, warm=1)
            print(f"lstm {prec} H={H} B={B}: {best:.2f} ms best, {B/best*1e3:.0f} windows/s, "
                  f"{B/best*1e3*557.8e6/1e12:.1f} TFLOP/s")
for n in (1 << 20, 1 << 24):
    sw = synth.make_ode_sweep(1, n)
    dev = {k: torch.tensor(v).cuda() for k, v in sw.items()}
    for S in (4, 8):
        for wt in (True, False):
            f = lambda: ode.solve_ensemble(n, p_open=dev["p_open"], p_closed=dev["p_closed"], rates=dev["rates"],
                                           alpha_arr=dev["alpha"], y0_mode="probs06", coupling=True, substeps=S, want_traj=wt)
            best, med = timeit(f)
            flop = 12 + 19 * S * 123 + 220
            print(f"ode rk4 n={n} S={S} traj={wt}: {best:.3f} ms, {n/best*1e3/1e9:.2f} G traj/s, {n/best*1e3*flop/1e12:.1f} TFLOP/s, "
                  f"write {n*240/best*1e3/1e9 if wt else 0:.0f} GB/s")
    f = lambda: ode.solve_ensemble(n, p_open=dev["p_open"], p_closed=dev["p_closed"], rates=dev["rates"],
                                   alpha_arr=dev["alpha"], y0_mode="probs06", coupling=True, mode="rk45")
    best, med = timeit(f, reps=3, warm=1)
    print(f"ode rk45 n={n}: {best:.3f} ms, {n/best*1e3/1e6:.1f} M traj/s")
