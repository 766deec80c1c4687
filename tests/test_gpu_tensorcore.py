"""GPU unit tests of the two tcgen05 kernels in isolation (diagnostic C-ABI entry points), then the
bf16 forward end to end against the fp32 oracle.

bf16-mode tolerance (north_star asks for it to be stated separately from the fp32 1e-5 bound):
operands, the projected input G and h_t are rounded to bf16 (8-bit mantissa) and sigma/tanh use
MUFU tanh.approx (rel. err ~5e-4); over 3 layers x 256 dependent steps this gives, on the golden
configuration, |dlogit| <= 3e-2 per unit of logit_gain-free head, |dP| <= 1e-2 and attention <= 2e-3.
The reference's own autocast-vs-fp32 gap at default init is 7.8e-4 / 2.4e-4."""
import ctypes as C

import numpy as np
import pytest
import torch

from lstm_ode_bci_b200 import _native as N
from lstm_ode_bci_b200 import lstm, synth
from oracle import torch_port

pytestmark = pytest.mark.gpu


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr())


# K = 512 / N % 256 != 0 select the 128-column W-block variant (H = 256 layers 1-2)
@pytest.mark.parametrize("M,Nn,K", [(128, 256, 64), (128, 256, 128), (1000, 1024, 128), (4173, 1024, 256), (77, 512, 256), (300, 2048, 64),
                                    (300, 2048, 512), (1000, 1024, 384), (129, 128, 256), (2500, 2048, 256),
                                    # M >= 512, N % 256 == 0: the CTA-pair kernel (lstm_bf16_gemm_pair.cu); odd / even numbers of
                                    # 128-row blocks, ragged last block, K = 64 ... 512, more tiles than clusters
                                    (4100, 2048, 512), (777, 256, 512), (512, 256, 64), (70000, 2048, 512), (33000, 256, 448)])
def test_proj_gemm_tcgen05_matches_matmul(M, Nn, K):
    g = torch.Generator(device="cuda").manual_seed(M + K)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.7).to(torch.bfloat16)
    W = (torch.randn(Nn, K, device="cuda", generator=g) * 0.2).to(torch.bfloat16)
    bias = torch.randn(Nn, device="cuda", generator=g)
    Cc = torch.full((M, Nn), float("nan"), device="cuda", dtype=torch.bfloat16)
    N.check(N.lib().bci_selftest_proj_gemm_bf16(_p(A), _p(W), _p(bias), _p(Cc), M, Nn, K, _stream()))
    torch.cuda.synchronize()
    want = A.float() @ W.float().T + bias
    got = Cc.float()
    assert torch.isfinite(got).all()
    err = (got - want).abs()
    tol = 2.0 ** -8 * want.abs() + 1e-3          # one bf16 rounding of the fp32 result
    assert bool((err <= tol).all()), float((err - tol).max())


@pytest.mark.parametrize("M,Nn,K", [(300, 1024, 128), (1000, 2048, 512), (4173, 2048, 256), (640, 256, 512), (70000, 2048, 512)])
def test_proj_gemm_blocked_layout_matches_matmul(M, Nn, K):
    """The same GEMMs writing the recurrence's streaming layout [row / 128][n / 8][row % 128][8] (one-CTA kernel below 512 rows,
    CTA pairs from there on): un-blocked on the host side and compared with the plain product."""
    g = torch.Generator(device="cuda").manual_seed(M + K + 1)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.7).to(torch.bfloat16)
    W = (torch.randn(Nn, K, device="cuda", generator=g) * 0.2).to(torch.bfloat16)
    bias = torch.randn(Nn, device="cuda", generator=g)
    mb = (M + 127) // 128
    Cb = torch.full((mb, Nn // 8, 128, 8), float("nan"), device="cuda", dtype=torch.bfloat16)
    N.check(N.lib().bci_selftest_proj_gemm_bf16_blocked(_p(A), _p(W), _p(bias), _p(Cb), M, Nn, K, _stream()))
    torch.cuda.synchronize()
    got = Cb.permute(0, 2, 1, 3).reshape(mb * 128, Nn)[:M].float()
    want = A.float() @ W.float().T + bias
    assert torch.isfinite(got).all()
    err = (got - want).abs()
    tol = 2.0 ** -8 * want.abs() + 1e-3
    assert bool((err <= tol).all()), float((err - tol).max())
    if M % 128:      # rows past M inside the last block hold bias only (zero-filled operand rows), never garbage from another tile
        pad = Cb.permute(0, 2, 1, 3).reshape(mb * 128, Nn)[M:].float()
        assert bool((pad - bias.to(torch.bfloat16).float()).abs().max() <= 1e-6)


def _perm(H=128, order="T"):
    """idx[natural row gate*H+unit] = permuted row; see include/bci_b200.h (perm_T / perm_G)."""
    idx = np.empty(4 * H, dtype=np.int64)
    for gate in range(4):
        for unit in range(H):
            half, slab, u = unit // 64, (unit // 8) % 8, unit % 8
            idx[gate * H + unit] = (half * 256 + slab * 32 if order == "T" else slab * 64 + half * 32) + gate * 8 + u
    return idx


@pytest.mark.parametrize("Bc,T", [(128, 4), (200, 9), (5, 33)])
def test_recurrence_tcgen05_matches_stepwise(Bc, T):
    H = 128
    g = torch.Generator(device="cuda").manual_seed(Bc * 131 + T)
    whh = [(torch.rand(4 * H, H, device="cuda", generator=g) * 2 - 1) / np.sqrt(H) for _ in range(2)]
    Gn = torch.randn(T, Bc, 2, 4 * H, device="cuda", generator=g) * 1.5       # natural gate order i,f,g,o
    perm = torch.from_numpy(_perm(H, "T")).cuda()
    perm_g = torch.from_numpy(_perm(H, "G")).cuda()
    # the kernel expects the 1/2 of sigmoid(x) = 0.5 + 0.5 tanh(x/2) folded into the i,f,o rows / columns
    gate_scale = torch.ones(4 * H, device="cuda")
    gate_scale[:2 * H] = 0.5
    gate_scale[3 * H:] = 0.5
    whh_p = []
    for d in range(2):
        wp = torch.empty_like(whh[d])
        wp[perm] = whh[d] * gate_scale[:, None]
        whh_p.append(wp.to(torch.bfloat16).contiguous())
    Gp = torch.empty_like(Gn)
    Gp[:, :, :, perm_g] = Gn * gate_scale
    Gp = Gp.reshape(T, Bc, 8 * H).to(torch.bfloat16).contiguous()
    # blocked streaming layout expected by the kernel: [row/128][dir*64 + chunk][row%128][8]
    M = T * Bc
    Mp = (M + 127) // 128 * 128
    flat = torch.zeros(Mp, 8 * H, device="cuda", dtype=torch.bfloat16)
    flat[:M] = Gp.reshape(M, 8 * H)
    Gblk = flat.reshape(Mp // 128, 128, 128, 8).permute(0, 2, 1, 3).contiguous()
    out = torch.full((T, Bc, 2 * H), float("nan"), device="cuda", dtype=torch.bfloat16)
    N.check(N.lib().bci_selftest_rec_bf16(_p(Gblk), _p(whh_p[0]), _p(whh_p[1]), _p(out), Bc, T, _stream()))
    torch.cuda.synchronize()
    # step-by-step emulation with the same roundings (bf16 G, bf16 weights, bf16 h fed back; fp32 c)
    Gq = Gp.float().reshape(T, Bc, 2, 4 * H)[:, :, :, perm_g] / gate_scale   # back to natural order and scale
    want = torch.empty(T, Bc, 2 * H, device="cuda")
    for d in range(2):
        w = (whh[d] * gate_scale[:, None]).to(torch.bfloat16).float() / gate_scale[:, None]
        h = torch.zeros(Bc, H, device="cuda")
        c = torch.zeros(Bc, H, device="cuda")
        for s in range(T):
            t = T - 1 - s if d else s
            pre = Gq[t, :, d] + h @ w.T
            i, f, gg, o = pre[:, :H].sigmoid(), pre[:, H:2 * H].sigmoid(), pre[:, 2 * H:3 * H].tanh(), pre[:, 3 * H:].sigmoid()
            c = f * c + i * gg
            hf = o * c.tanh()
            want[t, :, d * H:(d + 1) * H] = hf
            h = hf.to(torch.bfloat16).float()
    got = out.float()
    assert torch.isfinite(got).all()
    assert float((got - want).abs().max()) <= 1.5e-2      # bf16 output rounding (4e-3) + tanh.approx, compounding over T


# bf16 tensor-core mode vs the fp32 oracle at logit gain 12: 3x the measured errors (see the test below)
BF16_TOL = {"logits": 6e-3, "probs": 1.2e-3, "attn": 1.2e-4}


@pytest.mark.parametrize("B,T", [(8, 256), (130, 64), (37, 128)])
def test_bf16_forward_close_to_fp32_oracle(B, T):
    params = synth.make_lstm_params(42, 61, 128, 3, logit_gain=12.0)
    x = synth.make_windows(7, B, T, 61, structured=True)
    port = torch_port.build_port(params).eval()
    with torch.no_grad():
        want_logits, want_attn = port(torch.from_numpy(x), return_attention=True)
        want_probs = torch.softmax(want_logits, 1).numpy()
    m = lstm.from_params(params, precision="bf16")
    with torch.no_grad():
        logits, attn = m(torch.from_numpy(x).cuda(), return_attention=True)
        probs = m.predict_proba(torch.from_numpy(x).cuda())
    dl = np.abs(logits.cpu().numpy() - want_logits.numpy()).max()
    dp = np.abs(probs.cpu().numpy() - want_probs).max()
    da = np.abs(attn.cpu().numpy() - want_attn.numpy()).max()
    print(f"bf16 vs fp32 oracle: dlogit {dl:.3e} dprob {dp:.3e} dattn {da:.3e}")
    # stated bf16-mode tolerance = 3x the errors measured on this configuration (logit gain 12: logits 1.1-2.0e-3,
    # probabilities 1.5-3.9e-4, attention 1-4e-5): a regression of the recurrence's accuracy fails here
    assert dl <= BF16_TOL["logits"] and dp <= BF16_TOL["probs"] and da <= BF16_TOL["attn"], (dl, dp, da)
    # fp32 mode on the same inputs for reference of scale
    m32 = lstm.from_params(params, precision="fp32")
    with torch.no_grad():
        p32 = m32.predict_proba(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.abs(p32 - want_probs).max() <= 1e-5


@pytest.mark.parametrize("B,T", [(8, 256), (300, 64)])
def test_bf16_h256_forward_close_to_fp32_oracle(B, T):
    """hidden_size = 256 (the reference's trained checkpoint on 61 channels, 04:876-877) in bf16 mode: tcgen05 projection GEMM +
    the cluster recurrence of lstm_bf16_h256.cu; same stated bf16 tolerance as H = 128."""
    params = synth.make_lstm_params(44, 61, 256, 3, logit_gain=12.0)
    x = synth.make_windows(9, B, T, 61, structured=True)
    port = torch_port.build_port(params).eval()
    with torch.no_grad():
        want_logits, want_attn = port(torch.from_numpy(x), return_attention=True)
        want_probs = torch.softmax(want_logits, 1).numpy()
    m = lstm.from_params(params, precision="bf16")
    with torch.no_grad():
        logits, attn = m(torch.from_numpy(x).cuda(), return_attention=True)
        probs = m.predict_proba(torch.from_numpy(x).cuda())
    dl = np.abs(logits.cpu().numpy() - want_logits.numpy()).max()
    dp = np.abs(probs.cpu().numpy() - want_probs).max()
    da = np.abs(attn.cpu().numpy() - want_attn.numpy()).max()
    print(f"bf16 H=256 vs fp32 oracle: dlogit {dl:.3e} dprob {dp:.3e} dattn {da:.3e}")
    # measured at H = 256: logits 1.0-1.8e-3, probabilities 1.8-5.1e-4, attention 1-6e-5; asserted at 3x
    assert dl <= BF16_TOL["logits"] and dp <= 1.5e-3 and da <= 2e-4, (dl, dp, da)


def test_bf16_h256_full_chunk_matches_oracle_and_is_order_independent():
    """H = 256 bf16 at its own chunk size (`bci_lstm_chunk_windows`: 8 448 windows = 33 clusters x 256, two work items -- the two
    directions -- per cluster: the CTA-pair projection GEMM over 2.2 M rows and the staggered recurrence with barrier parities carried
    across items): reversing the window order reverses the outputs BIT FOR BIT, and three 16-window slices agree with the ORACLE
    (torch CPU port of the reference module at its checkpoint size) within the stated bf16 tolerance."""
    from lstm_ode_bci_b200 import ops
    params = synth.make_lstm_params(44, 61, 256, 3, logit_gain=12.0)
    m = lstm.from_params(params, precision="bf16")
    B = ops.lstm_chunk_windows(m._engine("bf16"))
    g = torch.Generator(device="cuda").manual_seed(23)
    x = torch.randn((B, 256, 61), device="cuda", generator=g)
    with torch.no_grad():
        p, a = m.predict_proba(x, return_attention=True)
        pr, ar = m.predict_proba(x.flip(0).contiguous(), return_attention=True)
    assert torch.isfinite(p).all() and torch.isfinite(a).all()
    assert torch.equal(pr.flip(0), p) and torch.equal(ar.flip(0), a)
    assert float((p.sum(1) - 1).abs().max()) <= 1e-6 and float((a.sum(1) - 1).abs().max()) <= 1e-5
    port = torch_port.build_port(params).eval()
    xc = x.cpu()
    for lo in (0, B // 2 - 8, B - 16):
        with torch.no_grad():
            wl, wa = port(xc[lo:lo + 16], return_attention=True)
            wp = torch.softmax(wl, 1)
        dp = float((p[lo:lo + 16].cpu() - wp).abs().max())
        da = float((a[lo:lo + 16].cpu() - wa).abs().max())
        print(f"H=256 full chunk vs oracle, windows {lo}..{lo + 16}: dprob {dp:.3e} dattn {da:.3e}")
        assert dp <= 1.5e-3 and da <= 2e-4, (lo, dp, da)


def test_bf16_rejects_ablation_variants_and_other_sizes():
    params = synth.make_lstm_params(1, 61, 128, 2, bidirectional=False)
    m = lstm.from_params(params, precision="bf16")
    with pytest.raises(N.BciError):
        m(torch.zeros(1, 8, 61, device="cuda"))


# (1, 1): one window, one step; (513, 2): 5 tiles = 2 tile quads, the second mostly empty; (20480, 3): 40 quads x 2 directions = 80
# work items on 33 resident clusters -> up to 3 items per cluster (persistent loop, weight reload between items)
@pytest.mark.parametrize("Bc,T,Kin", [(256, 3, 128), (256, 6, 256), (200, 9, 256), (5, 17, 128), (700, 5, 256), (1, 1, 256),
                                      (513, 2, 128), (20480, 3, 256)])
def test_fused_cluster_recurrence_matches_stepwise(Bc, T, Kin):
    """lstm_fused_bf16 (4-CTA cluster, cta_group::2 MMAs, h exchanged through DSMEM): projection + recurrence of one layer
    against a step-by-step emulation with the same roundings (bf16 inputs / weights / h fed back, fp32 gates and cell)."""
    H = 128
    g = torch.Generator(device="cuda").manual_seed(Bc * 7 + T * 3 + Kin)
    wih = [(torch.rand(4 * H, Kin, device="cuda", generator=g) * 2 - 1) / np.sqrt(H) for _ in range(2)]
    whh = [(torch.rand(4 * H, H, device="cuda", generator=g) * 2 - 1) / np.sqrt(H) for _ in range(2)]
    b = [(torch.rand(4 * H, device="cuda", generator=g) * 2 - 1) * 0.3 for _ in range(2)]
    x = (torch.randn(T, Bc, Kin, device="cuda", generator=g) * 0.8).to(torch.bfloat16).contiguous()
    perm = torch.from_numpy(_perm(H, "T")).cuda()
    gate_scale = torch.ones(4 * H, device="cuda")
    gate_scale[:2 * H] = 0.5
    gate_scale[3 * H:] = 0.5
    wih_p = torch.empty(2, 4 * H, Kin, device="cuda")
    whh_p, bias_p = [], torch.empty(2, 4 * H, device="cuda")
    for d in range(2):
        wih_p[d][perm] = wih[d] * gate_scale[:, None]
        wp = torch.empty_like(whh[d])
        wp[perm] = whh[d] * gate_scale[:, None]
        whh_p.append(wp.to(torch.bfloat16).contiguous())
        bias_p[d][perm] = b[d] * gate_scale
    wih_p = wih_p.to(torch.bfloat16).contiguous()
    bias_p = bias_p.contiguous()
    out = torch.full((T, Bc, 2 * H), float("nan"), device="cuda", dtype=torch.bfloat16)
    stats = torch.full((T, 8, Bc, 2), float("nan"), device="cuda")
    N.check(N.lib().bci_selftest_fused_rec_bf16(_p(x), _p(wih_p), _p(whh_p[0]), _p(whh_p[1]), _p(bias_p), _p(out), _p(stats),
                                                Bc, T, Kin, _stream()))
    torch.cuda.synchronize()
    want = torch.empty(T, Bc, 2 * H, device="cuda")
    for d in range(2):
        wi = (wih[d] * gate_scale[:, None]).to(torch.bfloat16).float() / gate_scale[:, None]
        wh = (whh[d] * gate_scale[:, None]).to(torch.bfloat16).float() / gate_scale[:, None]
        h = torch.zeros(Bc, H, device="cuda")
        c = torch.zeros(Bc, H, device="cuda")
        for s in range(T):
            t = T - 1 - s if d else s
            pre = x[t].float() @ wi.T + h @ wh.T + b[d]
            i, f, gg, o = pre[:, :H].sigmoid(), pre[:, H:2 * H].sigmoid(), pre[:, 2 * H:3 * H].tanh(), pre[:, 3 * H:].sigmoid()
            c = f * c + i * gg
            hf = o * c.tanh()
            want[t, :, d * H:(d + 1) * H] = hf
            h = hf.to(torch.bfloat16).float()
    got = out.float()
    assert torch.isfinite(got).all()
    assert float((got - want).abs().max()) <= 1.5e-2
    # LayerNorm partial statistics: 8 partials per row = [dir][32-unit group], stored slot-major [T][8][Bc]
    ssum = want.reshape(T, Bc, 8, 32).sum(-1).permute(0, 2, 1)
    ssq = (want.reshape(T, Bc, 8, 32) ** 2).sum(-1).permute(0, 2, 1)
    assert float((stats[..., 0] - ssum).abs().max()) <= 5e-2 and float((stats[..., 1] - ssq).abs().max()) <= 5e-2


def test_fused_kernel_is_robust_to_timing_jitter():
    """The cluster kernel's cross-CTA protocol (mbarriers, DSMEM copies, relays) must not depend on the natural timing:
    with BCI_FUSED_JITTER every role sleeps a pseudo-random time (up to 4 us) at its synchronisation points.  An earlier
    version re-armed an mbarrier too early and only failed (launch failure) under such perturbation or under ncu.  The same
    perturbation is applied to the CTA-pair recurrence of the fp32 path (lstm_rec_f16x3: commit multicast, h_local, peer relay)."""
    import os, subprocess, sys
    code = (
        "import numpy as np, torch\n"
        "from lstm_ode_bci_b200 import lstm, synth\n"
        "p = synth.make_lstm_params(42, 61, 128, 3, logit_gain=12.0)\n"
        "x = torch.from_numpy(synth.make_windows(7, 700, 256, 61, structured=True)).cuda()\n"
        "a = lstm.from_params(p, precision='bf16').predict_proba(x).cpu().numpy()\n"
        "b = lstm.from_params(p, precision='fp32').predict_proba(x[:64]).cpu().numpy()\n"
        "assert np.isfinite(a).all() and np.abs(a[:64] - b).max() <= 1e-2, np.abs(a[:64] - b).max()\n"
        "xf = torch.from_numpy(synth.make_windows(8, 2100, 24, 61, structured=True)).cuda()\n"   # 18 work items: the fp32 pair recurrence
        "m32 = lstm.from_params(p, precision='fp32')\n"
        "c = m32.predict_proba(xf).cpu().numpy()\n"
        "d = torch.cat([m32.predict_proba(xf[:1000]), m32.predict_proba(xf[1000:2000])]).cpu().numpy()\n"   # CUDA-core recurrence
        "assert np.isfinite(c).all() and np.abs(c[:2000] - d).max() <= 2e-6, np.abs(c[:2000] - d).max()\n"
        "print('jitter ok')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, BCI_FUSED_JITTER="4096", PYTHONPATH=root)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "jitter ok" in r.stdout, r.stderr[-2000:]


def test_staggered_h256_recurrence_is_robust_to_timing_jitter(tmp_path):
    """lstm_rec256_bf16_pipe exchanges K atoms between CTA pairs per phase (h_in / copy_done / st_free / recv_ready / a0_free per atom):
    under BCI_FUSED_JITTER (every role sleeps up to 4 us at its synchronisation points) the H = 256 bf16 forward must return the very bits
    of the unperturbed run -- 700 windows (three tile pairs per direction: several work items per cluster generation) x 40 steps."""
    import os, subprocess, sys
    code = (
        "import sys, numpy as np, torch\n"
        "from lstm_ode_bci_b200 import lstm, synth\n"
        "p = synth.make_lstm_params(44, 61, 256, 3, logit_gain=12.0)\n"
        "x = torch.from_numpy(synth.make_windows(9, 700, 40, 61, structured=True)).cuda()\n"
        "pr, at = lstm.from_params(p, precision='bf16').predict_proba(x, return_attention=True)\n"
        "assert torch.isfinite(pr).all() and torch.isfinite(at).all()\n"
        "np.savez(sys.argv[1], probs=pr.cpu().numpy(), attn=at.cpu().numpy())\n"
        "print('run ok')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for tag, jitter in (("plain", None), ("jitter", "4096")):
        env = dict(os.environ, PYTHONPATH=root)
        env.pop("BCI_FUSED_JITTER", None)
        if jitter:
            env["BCI_FUSED_JITTER"] = jitter
        f = str(tmp_path / (tag + ".npz"))
        r = subprocess.run([sys.executable, "-c", code, f], env=env, capture_output=True, text=True, timeout=240)
        assert r.returncode == 0 and "run ok" in r.stdout, r.stderr[-2000:]
        outs.append(np.load(f))
    assert np.array_equal(outs[0]["probs"], outs[1]["probs"]) and np.array_equal(outs[0]["attn"], outs[1]["attn"])


def _perm256():
    idx = np.empty(1024, dtype=np.int64)
    for gate in range(4):
        for unit in range(256):
            idx[gate * 256 + unit] = (unit // 128) * 512 + ((unit // 64) % 2) * 256 + ((unit // 8) % 8) * 32 + gate * 8 + unit % 8
    return idx


@pytest.mark.parametrize("Bc,T", [(256, 4), (200, 9), (5, 33), (700, 3)])
def test_recurrence_h256_cluster_matches_stepwise(Bc, T):
    """lstm_rec256_bf16 (4-CTA cluster, W_hh resident, h exchanged through DSMEM) against a step-by-step emulation with the
    same roundings (bf16 G, bf16 weights, bf16 h fed back; fp32 c)."""
    H = 256
    g = torch.Generator(device="cuda").manual_seed(Bc * 131 + T + 256)
    whh = [(torch.rand(4 * H, H, device="cuda", generator=g) * 2 - 1) / np.sqrt(H) for _ in range(2)]
    Gn = torch.randn(T, Bc, 2, 4 * H, device="cuda", generator=g) * 1.5
    perm = torch.from_numpy(_perm256()).cuda()
    gate_scale = torch.ones(4 * H, device="cuda")
    gate_scale[:2 * H] = 0.5
    gate_scale[3 * H:] = 0.5
    whh_p = torch.empty(2, 4 * H, H, device="cuda")
    for d in range(2):
        whh_p[d][perm] = whh[d] * gate_scale[:, None]
    whh_p = whh_p.to(torch.bfloat16).contiguous()
    Gp = torch.empty_like(Gn)
    Gp[:, :, :, perm] = Gn * gate_scale
    Gp = Gp.reshape(T, Bc, 8 * H).to(torch.bfloat16).contiguous()
    M = T * Bc
    Mp = (M + 127) // 128 * 128
    flat = torch.zeros(Mp, 8 * H, device="cuda", dtype=torch.bfloat16)
    flat[:M] = Gp.reshape(M, 8 * H)
    Gblk = flat.reshape(Mp // 128, 128, 256, 8).permute(0, 2, 1, 3).contiguous()
    out = torch.full((T, Bc, 2 * H), float("nan"), device="cuda", dtype=torch.bfloat16)
    N.check(N.lib().bci_selftest_rec256_bf16(_p(Gblk), _p(whh_p), _p(out), Bc, T, _stream()))
    torch.cuda.synchronize()
    Gq = Gp.float().reshape(T, Bc, 2, 4 * H)[:, :, :, perm] / gate_scale
    want = torch.empty(T, Bc, 2 * H, device="cuda")
    for d in range(2):
        w = (whh[d] * gate_scale[:, None]).to(torch.bfloat16).float() / gate_scale[:, None]
        h = torch.zeros(Bc, H, device="cuda")
        c = torch.zeros(Bc, H, device="cuda")
        for s in range(T):
            t = T - 1 - s if d else s
            pre = Gq[t, :, d] + h @ w.T
            i, f, gg, o = pre[:, :H].sigmoid(), pre[:, H:2 * H].sigmoid(), pre[:, 2 * H:3 * H].tanh(), pre[:, 3 * H:].sigmoid()
            c = f * c + i * gg
            hf = o * c.tanh()
            want[t, :, d * H:(d + 1) * H] = hf
            h = hf.to(torch.bfloat16).float()
    got = out.float()
    assert torch.isfinite(got).all()
    assert float((got - want).abs().max()) <= 1.5e-2


def _tf32x3(mode, A, B, bias, M, Nn, K):
    C_ = torch.full((M, Nn), float("nan"), device="cuda")
    N.check(N.lib().bci_selftest_gemm_tf32x3(mode, A.data_ptr(), B.data_ptr(), bias.data_ptr() if bias is not None else None,
                                             C_.data_ptr(), M, Nn, K, 0, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return C_


@pytest.mark.parametrize("M,Nn,K", [(128, 128, 32), (1000, 256, 128), (4097, 1024, 256), (3000, 128, 1024), (640, 160, 36)])
def test_gemm_tf32x3_nt_is_fp32_grade(M, Nn, K):
    """Split-precision tcgen05 GEMM of the fp32 path (projections, data gradients): C = A . B^T + bias against fp64;
    tolerance 4e-6 of max|C| (a cuBLAS fp32 GEMM measures 0.2-1.4e-6 on the same inputs, plain TF32 ~5e-4); ragged M,
    K tails and partial N blocks are handled by the tensor maps."""
    g = torch.Generator(device="cuda").manual_seed(M + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(Nn, K, device="cuda", generator=g) * 0.1
    bias = torch.randn(Nn, device="cuda", generator=g)
    ref = A.double() @ B.double().T + bias.double()
    got = _tf32x3(0, A, B, bias, M, Nn, K)
    assert not torch.isnan(got).any()
    assert float((got.double() - ref).abs().max() / ref.abs().max()) <= 4e-6


@pytest.mark.parametrize("P,Q,R", [(128, 128, 1024), (512, 128, 4000), (1024, 256, 16384)])
def test_gemm_tf32x3_tn_is_fp32_grade(P, Q, R):
    """Weight-gradient shape: C[P][Q] = sum_r A[r][P] B[r][Q] with MN-major operand tiles read straight from the row-major
    activations, split-K partial tiles combined by the TMA reduce-add (order-nondeterministic fp32 adds)."""
    g = torch.Generator(device="cuda").manual_seed(P + R)
    A = torch.randn(R, P, device="cuda", generator=g)
    B = torch.randn(R, Q, device="cuda", generator=g) * 0.1
    ref = A.double().T @ B.double()
    got = _tf32x3(1, A, B, None, P, Q, R)
    assert not torch.isnan(got).any()
    assert float((got.double() - ref).abs().max() / ref.abs().max()) <= 6e-6
    # one-hot operands: exact products land in exactly the right cells (layout / swizzle check)
    A0 = torch.zeros(R, P, device="cuda"); B0 = torch.zeros(R, Q, device="cuda")
    A0[9, 70] = 1.0; B0[9, 33] = 3.0; A0[R - 1, 1] = 1.0; B0[R - 1, Q - 2] = 5.0
    got = _tf32x3(1, A0, B0, None, P, Q, R)
    assert got.nonzero().tolist() == [[1, Q - 2], [70, 33]] and float(got[1, Q - 2]) == 5.0 and float(got[70, 33]) == 3.0


def test_bf16_single_pass_pooling_large_batch():
    """Batches of >= 4096 windows pool with the single-pass kernel (lstm_bf16_pool_stream.cu: score GEMM, softmax weights on the
    diagonal of a bf16 tile, context accumulated by a second MMA in TMEM).  4100 windows = 32 full 128-window blocks + a partial
    one; reference = the fp32 CUDA path (itself <= 1e-5 from the oracle) on the head, the tail and a middle slice."""
    B, T = 4100, 256
    params = synth.make_lstm_params(42, 61, 128, 3, logit_gain=12.0)
    x = torch.from_numpy(synth.make_windows(11, B, T, 61, structured=True)).cuda()
    mb = lstm.from_params(params, precision="bf16")
    m32 = lstm.from_params(params, precision="fp32")
    with torch.no_grad():
        pb, ab = mb.predict_proba(x, return_attention=True)
        pb2 = mb.predict_proba(x)
    assert torch.isfinite(pb).all() and torch.isfinite(ab).all()
    assert torch.equal(pb, pb2)                                   # attention output on/off does not change the probabilities
    assert float((ab.sum(dim=1) - 1).abs().max()) <= 1e-5         # softmax over T
    for sl in (slice(0, 96), slice(2000, 2064), slice(B - 100, B)):
        with torch.no_grad():
            p32, a32 = m32.predict_proba(x[sl], return_attention=True)
        assert float((pb[sl] - p32).abs().max()) <= BF16_TOL["probs"] and float((ab[sl] - a32).abs().max()) <= BF16_TOL["attn"]


def test_bf16_full_wave_properties():
    """Size-independent properties at the benchmark's full size (one full wave of the fused kernel, BASELINE configs[1]):
    windows are independent, so reversing their order reverses the outputs BIT FOR BIT (any cross-window leak in the cluster
    kernel's h exchange, the diag(beta) context MMA or the tile bookkeeping breaks this); probabilities and attention are
    normalised; a slice computed alone (other kernels: small-batch pooling) agrees within the bf16 tolerance."""
    params = synth.make_lstm_params(42, 61, 128, 3, logit_gain=12.0)
    m = lstm.from_params(params, precision="bf16")
    from lstm_ode_bci_b200 import ops
    B = ops.lstm_chunk_windows(m._engine("bf16"))
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((B, 256, 61), device="cuda", generator=g)
    with torch.no_grad():
        p, a = m.predict_proba(x, return_attention=True)
        pr, ar = m.predict_proba(x.flip(0).contiguous(), return_attention=True)
        ps = m.predict_proba(x[1000:1100].contiguous())
    assert torch.isfinite(p).all() and torch.isfinite(a).all()
    assert torch.equal(pr.flip(0), p) and torch.equal(ar.flip(0), a)
    assert float((p.sum(1) - 1).abs().max()) <= 1e-6 and float((a.sum(1) - 1).abs().max()) <= 1e-5
    assert float((ps - p[1000:1100]).abs().max()) <= BF16_TOL["probs"]
    # ... and against the ORACLE (torch CPU port of the reference module) on three 32-window slices of this very wave: the first
    # tile, a middle one that the persistent loop reaches as a later work item, and the last (the wave's final cluster)
    port = torch_port.build_port(params).eval()
    xc = x.cpu()
    for lo in (0, B // 2 - 16, B - 32):
        with torch.no_grad():
            wl, wa = port(xc[lo:lo + 32], return_attention=True)
            wp = torch.softmax(wl, 1)
        dp = float((p[lo:lo + 32].cpu() - wp).abs().max())
        da = float((a[lo:lo + 32].cpu() - wa).abs().max())
        print(f"full wave vs oracle, windows {lo}..{lo + 32}: dprob {dp:.3e} dattn {da:.3e}")
        assert dp <= BF16_TOL["probs"] and da <= BF16_TOL["attn"], (lo, dp, da)


def test_bf16_view_input_equals_packed_windows():
    """bci_lstm_forward_view: windows read in place from (R, S, C) recordings with 50 % overlap (02_preprocessing.py:157-180), in
    fp32 and in bf16, against the same windows materialised as the reference does -- BIT-identical in the bf16 mode (its input
    projection rounds x to bf16 on load, so storing bf16 loses nothing), and a pass that starts in the middle of a recording and
    crosses into the next one (first_window) returns the matching rows."""
    params = synth.make_lstm_params(42, 61, 128, 3, logit_gain=12.0)
    m = lstm.from_params(params, precision="bf16")
    R, S, C_, T, step = 3, 2000, 61, 256, 128
    g = torch.Generator(device="cuda").manual_seed(5)
    rec = torch.randn((R, S, C_), device="cuda", generator=g)
    n_seq = (S - T) // step + 1
    X = torch.stack([rec[r, i * step:i * step + T] for r in range(R) for i in range(n_seq)]).contiguous()
    with torch.no_grad():
        want, want_a = m.predict_proba(X, return_attention=True)
        got32, got32_a = m.predict_proba_recordings(rec, T, step, return_attention=True)
        got16 = m.predict_proba_recordings(rec.to(torch.bfloat16), T, step)
        part = m.predict_proba_recordings(rec.to(torch.bfloat16), T, step, first_window=n_seq - 3, n_windows=7)
    assert want.shape[0] == R * n_seq
    assert torch.equal(got32, want) and torch.equal(got32_a, want_a)
    assert torch.equal(got16, want)
    assert torch.equal(part, want[n_seq - 3:n_seq + 4])
    # fp32 parity mode through the same view (CUDA-core input projection): equal to its packed-window result as well
    m32 = lstm.from_params(params, precision="fp32")
    with torch.no_grad():
        assert torch.equal(m32.predict_proba_recordings(rec, T, step), m32.predict_proba(X))



@pytest.mark.parametrize("Bc,T,ND", [(256, 8, 2), (300, 33, 2), (5, 40, 1), (4300, 6, 2), (129, 64, 2)])
def test_recurrence_f16x3_is_fp32_grade(Bc, T, ND):
    """lstm_rec_f16x3 (CTA pair, cta_group::2, h . W_hh^T as three fp16 MMA chains: lo.hi + hi.lo + hi.hi) against a float64
    step-by-step recurrence: fp32-grade (the bf16 recurrence kernels are asserted at 1.5e-2 on the same quantity).  Covers a
    single pair, ragged tiles (300, 129: the pair's second CTA mostly / entirely dead), one direction, and more work items than
    resident clusters (4300 windows x 2 directions = 34 items)."""
    H = 128
    g = torch.Generator(device="cuda").manual_seed(Bc * 13 + T + ND)
    whh = (torch.rand(ND, 4 * H, H, device="cuda", generator=g) * 2 - 1) / np.sqrt(H) * 1.5
    G = (torch.randn(T * Bc, ND * 4 * H, device="cuda", generator=g) * 1.2).contiguous()     # column dir*512 + unit*4 + gate
    packed = torch.empty(ND * 2 * 4 * H * H, device="cuda", dtype=torch.float16)
    out = torch.full((T, Bc, ND * H), float("nan"), device="cuda")
    N.check(N.lib().bci_selftest_rec_f16x3(_p(G), _p(whh.contiguous()), _p(packed), _p(out), Bc, T, ND, _stream()))
    torch.cuda.synchronize()
    want = torch.empty(T, Bc, ND * H, device="cuda", dtype=torch.float64)
    G4 = G.double().reshape(T, Bc, ND, H, 4)                  # (unit, gate) interleaved
    for d in range(ND):
        w = whh[d].double()                                   # rows gate*H + unit
        h = torch.zeros(Bc, H, device="cuda", dtype=torch.float64)
        c = torch.zeros_like(h)
        for s in range(T):
            t = T - 1 - s if d else s
            rec = (h @ w.T).reshape(Bc, 4, H)                 # (gate, unit)
            pre = G4[t, :, d] + rec.permute(0, 2, 1)          # (unit, gate)
            i, f, gg, o = pre[..., 0].sigmoid(), pre[..., 1].sigmoid(), pre[..., 2].tanh(), pre[..., 3].sigmoid()
            c = f * c + i * gg
            h = o * c.tanh()
            want[t, :, d * H:(d + 1) * H] = h
    assert torch.isfinite(out).all()
    err = float((out.double() - want).abs().max())
    print(f"f16x3 recurrence vs float64: max abs err {err:.3e} (Bc={Bc}, T={T}, ND={ND})")
    assert err <= 3e-6



@pytest.mark.parametrize("M,Nn,K", [(128, 128, 64), (1000, 256, 128), (4097, 1024, 256), (640, 160, 72)])
def test_gemm_f16x3_nt_is_fp32_grade(M, Nn, K):
    """fp16-split form of the projection GEMM (A and B as fp16 (hi, lo) pairs, three kind::f16 MMA chains, B pre-scaled by 16):
    C = A . B^T + bias against fp64 on LSTM-like operands (activations O(1), weights ~1/sqrt(K)); same 4e-6-of-max|C| bound as
    the 3 x TF32 form; ragged M, a K tail and a partial N block."""
    g = torch.Generator(device="cuda").manual_seed(M + K + 1)
    A = torch.randn(M, K, device="cuda", generator=g) * 0.7
    B = (torch.rand(Nn, K, device="cuda", generator=g) * 2 - 1) / np.sqrt(K) * 1.5
    bias = torch.randn(Nn, device="cuda", generator=g)
    ref = A.double() @ B.double().T + bias.double()
    C_ = torch.full((M, Nn), float("nan"), device="cuda")
    N.check(N.lib().bci_selftest_gemm_f16x3(_p(A), _p(B), _p(bias), _p(C_), M, Nn, K, _stream()))
    torch.cuda.synchronize()
    assert not torch.isnan(C_).any()
    err = float((C_.double() - ref).abs().max() / ref.abs().max())
    print(f"f16x3 GEMM rel err {err:.2e} (M={M}, N={Nn}, K={K})")
    assert err <= 4e-6


def test_bf16_recording_view_at_the_reference_recording_shape_matches_oracle():
    """The end-to-end path of the bench at the reference's own recording shape: one 300 s x 500 Hz recording (150 000 samples x 61
    channels, 01:51-52) stored as bf16, its 1 170 overlapping windows (02:49-51,169) read in place by the input projection, against
    the ORACLE (torch CPU port of the reference module, fed the fp32 windows create_sequences would cut) on the first, a middle and
    the last 16 windows; stated bf16-mode tolerance."""
    S, C_, T, step = 150_000, 61, 256, 128
    n_seq = (S - T) // step + 1
    assert n_seq == 1170
    params = synth.make_lstm_params(42, 61, 128, 3, logit_gain=12.0)
    g = torch.Generator().manual_seed(11)
    t = torch.arange(S, dtype=torch.float32)[:, None] / 500.0
    freqs = torch.rand(1, C_, generator=g) * 40 + 2
    rec = (torch.sin(2 * np.pi * freqs * t + torch.rand(1, C_, generator=g) * 6.28) + 0.5 * torch.randn(S, C_, generator=g))
    rec = ((rec - rec.mean(0)) / rec.std(0))[None].contiguous()                  # z-scored per channel (02:134-154), (1, S, C)
    m = lstm.from_params(params, precision="bf16")
    with torch.no_grad():
        probs, attn = m.predict_proba_recordings(rec.to(torch.bfloat16).cuda(), T, step, return_attention=True)
    assert probs.shape == (n_seq, 2) and torch.isfinite(probs).all()
    port = torch_port.build_port(params).eval()
    for lo in (0, 577, n_seq - 16):
        X = torch.stack([rec[0, (lo + i) * step:(lo + i) * step + T] for i in range(16)])
        with torch.no_grad():
            wl, wa = port(X, return_attention=True)
            wp = torch.softmax(wl, 1)
        dp = float((probs[lo:lo + 16].cpu() - wp).abs().max())
        da = float((attn[lo:lo + 16].cpu() - wa).abs().max())
        print(f"recording view vs oracle, windows {lo}..{lo + 16}: dprob {dp:.3e} dattn {da:.3e}")
        assert dp <= BF16_TOL["probs"] and da <= BF16_TOL["attn"], (lo, dp, da)


@pytest.mark.parametrize("M,Nn,K,acc", [(512, 256, 64, 0), (1000, 1024, 256, 0), (4097, 512, 72, 0), (2048, 384, 128, 1), (131072, 1024, 128, 0),
                                        (300, 128, 64, 0)])
def test_gemm_tf32_single_pass_pair_kernel(M, Nn, K, acc):
    """Single-pass TF32 NT product of the mixed training step.  M >= 512 and N >= 256 run on CTA pairs (cta_group::2, one
    M256 x N256 x K8 MMA per K slice, each CTA loading its own A rows and half of the W rows); ragged M, a K tail, a partial N
    block, accumulation into C and the one-CTA fallback shape.  TF32 tolerance: 2e-3 of max|C| (10-bit mantissas, truncated operands)."""
    g = torch.Generator(device="cuda").manual_seed(M + K + Nn)
    A = torch.randn(M, K, device="cuda", generator=g) * 0.7
    B = (torch.rand(Nn, K, device="cuda", generator=g) * 2 - 1) / np.sqrt(K) * 1.5
    bias = torch.randn(Nn, device="cuda", generator=g)
    C0 = torch.randn(M, Nn, device="cuda", generator=g) if acc else None
    C_ = C0.clone() if acc else torch.full((M, Nn), float("nan"), device="cuda")
    ref = A.double() @ B.double().T + bias.double() + (C0.double() if acc else 0.0)
    N.check(N.lib().bci_selftest_gemm_tf32_single(_p(A), _p(B), _p(bias), _p(C_), M, Nn, K, acc, _stream()))
    torch.cuda.synchronize()
    assert not torch.isnan(C_).any()
    err = float((C_.double() - ref).abs().max() / ref.abs().max())
    print(f"single-pass TF32 GEMM rel err {err:.2e} (M={M}, N={Nn}, K={K}, accumulate={acc})")
    assert err <= 2e-3


@pytest.mark.parametrize("M,Nn,K", [(256, 128, 1024), (1024, 256, 131072), (512, 128, 130560), (1024, 128, 5000), (128, 256, 4096)])
def test_gemm_tf32_single_pass_tn_pair_kernel(M, Nn, K):
    """Single-pass TF32 TN product (weight gradients of the mixed step): C[M][N] = A[K][M]^T . B[K][N], MN-major operands read
    straight from the row-major activations, split-K with reduce-add.  M % 256 == 0 runs on CTA pairs (N = 256: 256 x 256 tiles,
    N = 128: 256 x 128); K with a tail block; the last shape is the one-CTA fallback."""
    g = torch.Generator(device="cuda").manual_seed(M + K + Nn)
    A = torch.randn(K, M, device="cuda", generator=g) * 0.5
    B = torch.randn(K, Nn, device="cuda", generator=g) * 0.5
    C_ = torch.full((M, Nn), float("nan"), device="cuda")
    N.check(N.lib().bci_selftest_gemm_tf32_single_tn(_p(A), _p(B), _p(C_), M, Nn, K, _stream()))
    torch.cuda.synchronize()
    ref = A.double().T @ B.double()
    assert not torch.isnan(C_).any()
    err = float((C_.double() - ref).abs().max() / ref.abs().max())
    print(f"single-pass TF32 TN GEMM rel err {err:.2e} (M={M}, N={Nn}, K={K})")
    assert err <= 3e-3
