"""Single-pass TF32 NT GEMM of the mixed training step at its own shapes: TFLOP/s and bytes/s (BCI_GEMM_PAIR=off: one-CTA kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lstm_ode_bci_b200 import _native as N
lib = N.lib()
st = torch.cuda.current_stream().cuda_stream
for (M, Nn, K) in ((131072, 1024, 128), (131072, 1024, 256), (131072, 256, 1024), (131072, 2048, 256), (131072, 2048, 512), (131072, 512, 2048)):
    A = torch.randn(M, K, device="cuda"); B = torch.randn(Nn, K, device="cuda") * 0.05; bias = torch.zeros(Nn, device="cuda")
    C = torch.empty(M, Nn, device="cuda")
    f = lambda: N.check(lib.bci_selftest_gemm_tf32_single(A.data_ptr(), B.data_ptr(), bias.data_ptr(), C.data_ptr(), M, Nn, K, 0, st))
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"M={M} N={Nn} K={K}: {ms * 1e3:.0f} us  {2.0 * M * Nn * K / ms / 1e9:.0f} TFLOP/s  HBM (A + C once) {(M * K + M * Nn) * 4 / ms / 1e9:.2f} TB/s")
