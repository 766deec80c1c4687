"""bci_preprocess on 36 raw fp32 recordings (61 x 150 000): CUDA-event timing of the whole call -- profiling target for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lstm_ode_bci_b200 import preprocessing as pp

R = int(sys.argv[1]) if len(sys.argv) > 1 else 36
raw = torch.randn((R, 61, 150000), device="cuda") * 1e-5 + 1e-4
b_, a_, zi_, padlen = pp.design_bandpass()
for _ in range(2):
    out = pp.preprocess_recordings(raw, b_, a_, zi_, padlen)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
torch.cuda.synchronize()
ev[0].record()
for _ in range(5):
    out = pp.preprocess_recordings(raw, b_, a_, zi_, padlen)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 5
print("bci_preprocess: %d recordings -> %d windows in %.3f ms (%.2f M windows/s)" % (R, out["X"].shape[0], ms, out["X"].shape[0] / ms / 1e3))
