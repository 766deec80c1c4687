"""One bf16 forward of the hidden-size-256 model over one chunk (profiling target for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lstm_ode_bci_b200 import lstm, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8448
m = lstm.from_params(synth.make_lstm_params(44, 61, 256, 3), precision="bf16")
x = torch.randn(B, 256, 61, device="cuda")
for _ in range(2):
    p = m.predict_proba(x)
torch.cuda.synchronize()
print("ok", float(p.sum()))
