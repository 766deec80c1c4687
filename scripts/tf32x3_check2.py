import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lstm_ode_bci_b200 import _native as N
lib = N.lib()
st = lambda: torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
def run(mode, A, B, bias, C, M, Nn, K, explicit=0):
    N.check(lib.bci_selftest_gemm_tf32x3(mode, A.data_ptr(), B.data_ptr(), bias.data_ptr() if bias is not None else None,
                                         C.data_ptr(), M, Nn, K, explicit, st()))
    torch.cuda.synchronize()
    return C
M, Nn, K = 1000, 256, 128
A = torch.randn(M, K, device="cuda"); B = torch.randn(Nn, K, device="cuda"); C0 = torch.randn(M, Nn, device="cuda")
C = run(2, A, B, None, C0.clone(), M, Nn, K)
ref = C0.double() + A.double() @ B.double().T
print("NT accumulate (reduce-add): rel err %.3e" % float((C.double() - ref).abs().max() / ref.abs().max()))
for (P, Q, R) in ((128, 128, 1024), (256, 256, 4096)):
    A = torch.randn(R, P, device="cuda"); B = torch.randn(R, Q, device="cuda")
    ref = A.double().T @ B.double()
    for mode in (3, 1):
        C = run(mode, A, B, None, torch.full((P, Q), float("nan"), device="cuda"), P, Q, R)
        d = (C.double() - ref).abs()
        print("TN mode %d P%d Q%d R%d: rel err %.3e  |C|max %.3e nan %d  C[0,:4]=%s ref[0,:4]=%s" % (mode, P, Q, R, float(d.max() / ref.abs().max()), float(C.abs().nan_to_num().max()), int(torch.isnan(C).sum()), C[0, :4].tolist(), ref[0, :4].tolist()))
# structured probe for the MN-major layout: A = one-hot rows
P, Q, R = 128, 128, 1024
A = torch.zeros(R, P, device="cuda"); B = torch.zeros(R, Q, device="cuda")
A[3, 5] = 1.0; B[3, 70] = 2.0
C = run(3, A, B, None, torch.zeros(P, Q, device="cuda"), P, Q, R)
nz = C.nonzero().tolist()
print("one-hot probe: expected [[5, 70]] value 2 ->", nz[:8], [float(C[i, j]) for i, j in nz[:8]])

# second probe: k row 9 (second 8-row group), MN group 2
A = torch.zeros(R, P, device="cuda"); B = torch.zeros(R, Q, device="cuda")
A[9, 70] = 1.0; B[9, 33] = 3.0; A[40, 1] = 1.0; B[40, 2] = 5.0
C = run(3, A, B, None, torch.zeros(P, Q, device="cuda"), P, Q, R)
nz = C.nonzero().tolist()
print("probe 2: expected [[1,2]]=5, [[70,33]]=3 ->", nz[:8], [float(C[i, j]) for i, j in nz[:8]])
