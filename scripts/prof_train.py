"""Two training steps at B=512 (profiling target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lstm_ode_bci_b200 import lstm, synth, train
H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
m = lstm.from_params(synth.make_lstm_params(42, 61, H, 3), precision="fp32", dropout=0.4).train()
tr = train.FusedTrainer(m, class_weight=[0.8, 1.2])
x = torch.randn(512, 256, 61, device="cuda"); y = torch.arange(512, device="cuda") % 2
for i in range(2):
    loss, norm = tr.step(x, y, seed=i)
torch.cuda.synchronize(); print("ok", float(loss), float(norm))
