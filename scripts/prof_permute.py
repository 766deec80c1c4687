"""bci_permute_channels at the permutation-importance defaults of 07_explainability.py:287 (1 000 resident test windows, variants
gathered two recurrence waves at a time): CUDA-event timing of the gather alone (fp32 and bf16 output) -- profiling target for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from lstm_ode_bci_b200 import ops

n, T, C, rows = 1000, 256, 61, 2 * 16896
V = -(-rows // n)
rng = np.random.default_rng(0)
x = torch.randn(n, T, C, device="cuda")
perm = torch.from_numpy(np.stack([rng.permutation(n) for _ in range(V)]).astype(np.int32)).cuda().view(-1)
ch = torch.from_numpy((np.arange(V) % C).astype(np.int32)).cuda()
for bf16 in (False, True):
    for _ in range(3):
        out = ops.permute_channels(x, perm, ch, 0, rows, bf16_out=bf16)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(10):
        out = ops.permute_channels(x, perm, ch, 0, rows, bf16_out=bf16)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 10
    wbytes = rows * T * C * (2 if bf16 else 4)
    print("permute_channels %s: %.3f ms per %d rows, write %.2f GB -> %.0f GB/s written (+ %.0f GB/s of L2-resident reads)"
          % ("bf16" if bf16 else "fp32", ms, rows, wbytes / 1e9, wbytes / ms / 1e6, rows * T * C * 4 / ms / 1e6))
