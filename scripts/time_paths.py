"""Device-resident timing of the bf16 forward for a given batch (BCI_BF16_PATH=split|fused picks the recurrence path)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lstm_ode_bci_b200 import lstm, synth, ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16896
params = synth.make_lstm_params(42, 61, 128, 3, logit_gain=12.0)
m = lstm.from_params(params, precision="bf16")
x = torch.randn(B, 256, 61, device="cuda")
hid = m._engine("bf16")
ops.lstm_set_profiling(hid, True)
for _ in range(3):
    m.predict_proba(x)
torch.cuda.synchronize()
ops.lstm_get_profile(hid)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
n = 5
for _ in range(n):
    p = m.predict_proba(x)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / n
prof = ops.lstm_get_profile(hid)
print("B", B, "ms/step %.3f" % ms, "windows/s %.0f" % (B / ms * 1e3), {k: round(v[0] / n, 3) for k, v in prof.items()})
# parity against the fp32 path on a subset
m32 = lstm.from_params(params, precision="fp32")
with torch.no_grad():
    p32 = m32.predict_proba(x[:256]).cpu().numpy()
print("dprob vs fp32 path", float(np.abs(p[:256].cpu().numpy() - p32).max()))
# attention and the tail of the batch (partial last block when B % 128 != 0)
with torch.no_grad():
    pb, ab = m.predict_proba(x, return_attention=True)
    p32t, a32t = m32.predict_proba(x[-200:], return_attention=True)
print("tail: dprob %.3e dattn %.3e  attn row sums %.6f..%.6f  nan %d" % (
    float((pb[-200:] - p32t).abs().max()), float((ab[-200:] - a32t).abs().max()), float(ab.sum(1).min()), float(ab.sum(1).max()),
    int(torch.isnan(pb).sum() + torch.isnan(ab).sum())))
