"""Small invocations of the cluster / DSMEM / mbarrier kernels for compute-sanitizer (racecheck, synccheck, memcheck):
lstm_fused_bf16 (4-CTA clusters, cta_group::2, DSMEM h exchange), lstm_rec256_bf16, attn_pool_stream_bf16, lstm_rec_f16x3
(CTA pairs), input_proj_bf16, gemm_tf32x3 / fp16-split GEMM.  Sizes are tiny: the tools slow kernels down by 10-100x.

    compute-sanitizer --tool racecheck python scripts/sanitize.py [which]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lstm_ode_bci_b200 import lstm, synth
which = sys.argv[1] if len(sys.argv) > 1 else "all"
g = torch.Generator(device="cuda").manual_seed(1)
if which in ("all", "fused"):
    m = lstm.from_params(synth.make_lstm_params(42, 61, 128, 3), precision="bf16")
    x = torch.randn(600, 128, 61, device="cuda", generator=g)       # 5 tiles -> 2 tile quads x 2 directions, tcgen05 K1 (T = 128)
    p = m.predict_proba(x); torch.cuda.synchronize(); print("fused bf16 ok", float(p.sum()))
if which in ("all", "pool"):
    m = lstm.from_params(synth.make_lstm_params(42, 61, 128, 1), precision="bf16")
    x = torch.randn(4224, 256, 61, device="cuda", generator=g)      # >= 4096 windows x 256 steps: single-pass pooling kernel
    p = m.predict_proba(x); torch.cuda.synchronize(); print("pool stream ok", float(p.sum()))
if which in ("all", "h256"):
    m = lstm.from_params(synth.make_lstm_params(44, 61, 256, 1), precision="bf16")
    x = torch.randn(300, 6, 61, device="cuda", generator=g)
    p = m.predict_proba(x); torch.cuda.synchronize(); print("h256 bf16 ok", float(p.sum()))
if which in ("all", "fp32tc"):
    m = lstm.from_params(synth.make_lstm_params(42, 61, 128, 1), precision="fp32")
    x = torch.randn(2100, 6, 61, device="cuda", generator=g)        # 9 pairs x 2 directions = 18 work items: pair recurrence + fp16-split GEMMs
    p = m.predict_proba(x); torch.cuda.synchronize(); print("fp32 tc ok", float(p.sum()))
