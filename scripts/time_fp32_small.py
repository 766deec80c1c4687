"""Inference at small batches (the reference's 256 / 512-window passes, single windows): ms per call.
Usage: python scripts/time_fp32_small.py [fp32|bf16]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lstm_ode_bci_b200 import lstm, synth

params = synth.make_lstm_params(42, 61, 128, 3, logit_gain=4.0)
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
m = lstm.from_params(params, precision=prec)
for B in (1, 32, 256, 512, 1024, 1536, 2048):
    x = torch.from_numpy(synth.make_windows(3, B, 256, 61)).cuda()
    with torch.no_grad():
        for _ in range(3):
            m.predict_proba(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            m.predict_proba(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{prec} forward B={B}: {ms:.3f} ms  ({B / ms:.1f} k windows/s)")
