"""N config-3 training steps (512 x 256, dropout 0.4) in one mode, for ncu launch lists: python scripts/train_step_once.py mixed|fp32 [steps] [H]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lstm_ode_bci_b200 import lstm, synth, train

mode = sys.argv[1] if len(sys.argv) > 1 else "mixed"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
H = int(sys.argv[3]) if len(sys.argv) > 3 else 128
params = synth.make_lstm_params(42, 61, H, 3, logit_gain=4.0)
x = torch.from_numpy(synth.make_windows(3, 512, 256, 61)).cuda()
y = (torch.arange(512) % 2).cuda()
m = lstm.from_params(params, precision="fp32", dropout=0.4).train()
tr = train.FusedTrainer(m, precision=mode)
for i in range(steps):
    loss, norm = tr.step(x, y, seed=i)
torch.cuda.synchronize()
print("loss", float(loss), "norm", float(norm))
