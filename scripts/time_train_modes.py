"""Config-3 training step (512 windows x 256 steps, H = 128) in the fp32-parity and the mixed mode, plus the swapped tensor-core
recurrences in isolation: CUDA-event timings on one GPU.  Usage: python scripts/time_train_modes.py [steps]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from lstm_ode_bci_b200 import _native as N
from lstm_ode_bci_b200 import lstm, synth, train


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    H, B, T = 256, 512, 256
    params = synth.make_lstm_params(42, 61, H, 3, logit_gain=4.0)
    x = torch.from_numpy(synth.make_windows(3, B, T, 61)).cuda()
    y = (torch.arange(B) % 2).cuda()
    for mode in ("fp32", "mixed"):
        m = lstm.from_params(params, precision="fp32", dropout=0.4).train()
        tr = train.FusedTrainer(m, precision=mode)
        ms = timed(lambda: tr.step(x, y, seed=1), max(3, steps // 2))
        print(f"H=256 train step {mode}: {ms:.3f} ms  ({B / ms:.1f} k windows/s)")
        del tr, m
    H, B, T = 128, 512, 256
    params = synth.make_lstm_params(42, 61, H, 3, logit_gain=4.0)
    x = torch.from_numpy(synth.make_windows(3, B, T, 61)).cuda()
    y = (torch.arange(B) % 2).cuda()
    for mode in ("fp32", "mixed"):
        m = lstm.from_params(params, precision="fp32", dropout=0.4).train()
        tr = train.FusedTrainer(m, precision=mode)
        ms = timed(lambda: tr.step(x, y, seed=1), steps)
        print(f"train step {mode}: {ms:.3f} ms  ({B / ms:.1f} k windows/s)")
    # isolated recurrences
    ND = 2
    g = torch.Generator(device="cuda").manual_seed(1)
    whh = ((torch.rand(ND, 4 * H, H, device="cuda", generator=g) * 2 - 1) / np.sqrt(H)).contiguous()
    G = torch.randn(T * B, ND * 4 * H, device="cuda", generator=g).contiguous()
    packed = torch.empty(5 * ND * 4 * H * H, device="cuda", dtype=torch.float16)
    out = torch.empty((T, B, ND * H), device="cuda")
    gates = torch.empty((T * B, ND * 4 * H), device="cuda")
    cs = torch.empty((T, B, ND * H), device="cuda")
    dG = torch.empty((T * B, ND * 4 * H), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    lib = N.lib()
    stamps = torch.zeros(32, dtype=torch.int64, device="cuda")
    lib.bci_selftest_swap_set_debug(stamps.data_ptr())
    N.check(lib.bci_selftest_rec_swap_fwd(G.data_ptr(), whh.data_ptr(), packed.data_ptr(), out.data_ptr(), gates.data_ptr(), cs.data_ptr(), B, T, ND, 0, st))
    torch.cuda.synchronize()
    lib.bci_selftest_swap_set_debug(None)
    sv = stamps.cpu().numpy().reshape(4, 8)
    print("forward timeline of CTA (0,0), SM cycles per phase [G prefetch, acc wait, tmem ld, epilogue, fence+arrive | (6,7) = MMA warp issue start, after commit] and step:")
    for r in range(4):
        d = np.diff(sv[r])
        nxt = (sv[r + 1][0] - sv[r][0]) if r < 3 else 0
        print("   ", d.tolist(), "step", int(nxt))
    for split in (0, 1):
        ms = timed(lambda: N.check(lib.bci_selftest_rec_swap_fwd(G.data_ptr(), whh.data_ptr(), packed.data_ptr(), out.data_ptr(), gates.data_ptr(), cs.data_ptr(), B, T, ND, split, st)), steps)
        print(f"swap forward recurrence split={split} (incl. 4 pack launches): {ms:.3f} ms per layer = {ms * 1e3 / T:.2f} us per step")
    dout = torch.randn_like(out) * 1e-3
    for split in (0, 1):
        ms = timed(lambda: N.check(lib.bci_selftest_bptt_swap(dout.data_ptr(), gates.data_ptr(), cs.data_ptr(), whh.data_ptr(), packed.data_ptr(), dG.data_ptr(), B, T, ND, split, st)), steps)
        print(f"swap BPTT recurrence split={split} (incl. 4 pack launches): {ms:.3f} ms per layer = {ms * 1e3 / T:.2f} us per step")


if __name__ == "__main__":
    main()
