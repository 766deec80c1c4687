"""Accuracy and speed of the split-precision tcgen05 GEMMs (csrc/gemm_tf32x3.cu) against fp64 / fp32 torch matmuls."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lstm_ode_bci_b200 import _native as N

lib = N.lib()
st = lambda: torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
torch.backends.cuda.matmul.allow_tf32 = False


def run(mode, A, B, bias, M, Nn, K, explicit):
    C = torch.full((M, Nn), float("nan"), device="cuda")
    N.check(lib.bci_selftest_gemm_tf32x3(mode, A.data_ptr(), B.data_ptr(), bias.data_ptr() if bias is not None else None,
                                         C.data_ptr(), M, Nn, K, explicit, st()))
    torch.cuda.synchronize()
    return C


for (M, Nn, K) in ((256, 128, 64), (1000, 256, 128), (4096, 1024, 256), (131072, 1024, 256), (131072, 256, 1024)):
    A = torch.randn(M, K, device="cuda"); B = torch.randn(Nn, K, device="cuda") * 0.1; bias = torch.randn(Nn, device="cuda")
    ref = (A.double() @ B.double().T + bias.double())
    e32 = float(((A @ B.T + bias).double() - ref).abs().max() / ref.abs().max())
    for ex in (0, 1):
        C = run(0, A, B, bias, M, Nn, K, ex)
        print("NT M%d N%d K%d explicit_hi=%d rel err %.3e (fp32 matmul %.3e) nan=%d" % (M, Nn, K, ex, float((C.double() - ref).abs().max() / ref.abs().max()), e32, int(torch.isnan(C).sum())), flush=True)
for (P, Q, R) in ((128, 128, 1024), (1024, 256, 131072), (1024, 128, 131072), (256, 128, 5000)):
    A = torch.randn(R, P, device="cuda"); B = torch.randn(R, Q, device="cuda") * 0.1
    ref = A.double().T @ B.double()
    e32 = float(((A.T @ B).double() - ref).abs().max() / ref.abs().max())
    for ex in (0, 1):
        C = run(1, A, B, None, P, Q, R, ex)
        print("TN P%d Q%d R%d explicit_hi=%d rel err %.3e (fp32 matmul %.3e) nan=%d" % (P, Q, R, ex, float((C.double() - ref).abs().max() / ref.abs().max()), e32, int(torch.isnan(C).sum())), flush=True)
# timing through the selftest entry (includes the split pass and a cudaMalloc: upper bound) vs torch fp32
import time
M, Nn, K = 131072, 1024, 256
A = torch.randn(M, K, device="cuda"); B = torch.randn(Nn, K, device="cuda"); bias = torch.randn(Nn, device="cuda")
for name, f in (("tf32x3 selftest", lambda: run(0, A, B, bias, M, Nn, K, 0)), ("torch fp32", lambda: A @ B.T + bias)):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): f()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print("%s: %.3f ms = %.1f TFLOP/s" % (name, dt * 1e3, 2.0 * M * Nn * K / dt / 1e12))
