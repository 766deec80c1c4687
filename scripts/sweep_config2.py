"""BASELINE configs[1]: inference sweep, batch 256 ... 65536 windows on one B200, fp32 parity mode and bf16 tensor-core mode.
Device-resident inputs, CUDA-event timing, >= 3 warm-ups, best of 5 and median; writes a JSON table (profiles/r1_sweep_config2.json, r2_sweep_config2.json)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lstm_ode_bci_b200 import lstm, synth

FLOP = 557_793_536
params = synth.make_lstm_params(42, 61, 128, 3)
out = {"flop_per_window": FLOP, "rows": []}
gen = torch.Generator(device="cuda").manual_seed(42)
xmax = torch.randn((65536, 256, 61), device="cuda", generator=gen)
for prec in ("bf16", "fp32"):
    m = lstm.from_params(params, precision=prec)
    for B in (1, 256, 512, 1024, 4096, 9472, 16384, 16896, 65536):
        x = xmax[:B]
        reps = 5 if (prec == "bf16" or B <= 16896) else 2
        for _ in range(3 if B <= 16896 else 1):
            p = m.predict_proba(x)
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); p = m.predict_proba(x); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        best, med = min(ts), float(np.median(ts))
        row = {"precision": prec, "windows": B, "ms_best": best, "ms_median": med, "windows_per_s": B / best * 1e3,
               "tflops": B * FLOP / best / 1e9, "p_open_mean": float(p[:, 0].mean())}
        out["rows"].append(row)
        print(row, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweep_config2.json", "w"), indent=1)
