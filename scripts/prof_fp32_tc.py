"""one fp32 inference forward at a full wave of the pair recurrence (for ncu)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lstm_ode_bci_b200 import lstm, ops, synth
m = lstm.from_params(synth.make_lstm_params(42, 61, 128, 3), precision="fp32")
B = ops.lstm_chunk_windows(m._engine("fp32"))
x = torch.randn(B, 256, 61, device="cuda")
with torch.no_grad():
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
        m.predict_proba(x)
torch.cuda.synchronize()
print("done", B)
