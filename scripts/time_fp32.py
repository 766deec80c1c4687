"""fp32 path timing: training step (512 windows) and inference forward (2048 windows), H from argv.  BCI_FP32_GEMM=simt selects
the CUDA-core GEMMs; default = split-precision tcgen05 GEMMs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lstm_ode_bci_b200 import lstm, synth, train
H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
params = synth.make_lstm_params(42, 61, H, 3)
m = lstm.from_params(params, precision="fp32", dropout=0.4).train()
tr = train.FusedTrainer(m, class_weight=[0.8, 1.2])
x = torch.randn(512, 256, 61, device="cuda"); y = torch.arange(512, device="cuda") % 2
for i in range(3):
    loss, norm = tr.step(x, y, seed=i)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(10):
    loss, norm = tr.step(x, y, seed=10 + i)
b.record(); torch.cuda.synchronize()
print("H=%d gemm=%s train step: %.2f ms (loss %.4f norm %.4f)" % (H, os.environ.get("BCI_FP32_GEMM", "tf32x3"), a.elapsed_time(b) / 10, float(loss), float(norm)))
m.eval()
B = 2048 if H == 128 else 1024
xi = torch.randn(B, 256, 61, device="cuda")
with torch.no_grad():
    for _ in range(2):
        p = m.predict_proba(xi)
    torch.cuda.synchronize(); a.record()
    for _ in range(5):
        p = m.predict_proba(xi)
    b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print("H=%d gemm=%s fp32 forward B=%d: %.2f ms = %.0f windows/s" % (H, os.environ.get("BCI_FP32_GEMM", "tf32x3"), B, ms, B / ms * 1e3))
