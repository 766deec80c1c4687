"""GPU checks of the mixed-precision training step (BCI_TRAIN_MIXED): the swapped tensor-core recurrences of
csrc/lstm_rec_swap.cu in isolation against float64 recurrences / torch autograd, and the whole step against the fp32-parity step
and the CPU port of the reference module (oracle/torch_port.py).

Stated tolerances of the mixed mode (forward operands fp16, BPTT operands bf16, single-pass TF32 GEMMs; the reference's own GPU
training runs under autocast fp16, 04_lstm_model.py:486-490): h_t within 2e-3 of float64; dG within 2 % of max|dG|; loss within
2e-3 of the fp32 reference; every gradient tensor at cosine >= 0.999 and within 5 % of its max-abs."""
import numpy as np
import pytest
import torch

from lstm_ode_bci_b200 import _native as N
from lstm_ode_bci_b200 import lstm, synth, train
from oracle import torch_port

pytestmark = pytest.mark.gpu


def _p(t):
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ref_forward64(G, whh, Bc, T, ND, H=128):
    """float64 step-by-step recurrence; G [T*Bc][ND*4H] (column dir*4H + unit*4 + gate), whh [ND][4H][H] PyTorch rows.
    Returns out [T][Bc][ND*H], gates [T][Bc][ND][H][4], c [T][Bc][ND][H] (differentiable wrt G if G requires grad)."""
    G4 = G.reshape(T, Bc, ND, H, 4)
    outs = [[None] * ND for _ in range(T)]
    gates = [[None] * ND for _ in range(T)]
    cs = [[None] * ND for _ in range(T)]
    for d in range(ND):
        w = whh[d]
        h = torch.zeros(Bc, H, device=G.device, dtype=G.dtype)
        c = torch.zeros_like(h)
        for s in range(T):
            t = T - 1 - s if d else s
            rec = (h @ w.T).reshape(Bc, 4, H)
            pre = G4[t, :, d] + rec.permute(0, 2, 1)
            i, f, gg, o = pre[..., 0].sigmoid(), pre[..., 1].sigmoid(), pre[..., 2].tanh(), pre[..., 3].sigmoid()
            c = f * c + i * gg
            h = o * c.tanh()
            outs[t][d], cs[t][d] = h, c
            gates[t][d] = torch.stack([i, f, gg, o], dim=-1)
    out = torch.stack([torch.cat(r, dim=1) for r in outs])
    gt = torch.stack([torch.stack(r, dim=1) for r in gates])
    ct = torch.stack([torch.stack(r, dim=1) for r in cs])
    return out, gt, ct


def test_tmem_a_operand_layout_probe():
    """tcgen05.mma with its A operand in tensor memory: lane = row, 32-bit column c = (K element 2c | K element 2c+1 << 16)."""
    out = torch.full((128, 16), float("nan"), device="cuda")
    N.check(N.lib().bci_selftest_tmem_a_probe(_p(out), _stream()))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    print("TMEM-A probe rows 0, 5, 127:", got[0], got[5], got[127])
    want = np.tile(np.arange(1, 17, dtype=np.float32), (128, 1))
    want[5, 0::2] += 100.0
    assert np.array_equal(got, want)


@pytest.mark.parametrize("split", [0, 1])
@pytest.mark.parametrize("Bc,T,ND", [(8, 8, 2), (512, 24, 2), (13, 40, 1), (300, 9, 2), (64, 256, 2)])
def test_rec_swap_forward_matches_float64(Bc, T, ND, split):
    """split=0: the mixed mode's fp16 product (one chain, tanh.approx gates); split=1: the fp32-parity form (three chains: lo.hi +
    hi.lo + hi.hi, ex2-based gates) -- fp32-grade like lstm_rec_f16x3."""
    H = 128
    g = torch.Generator(device="cuda").manual_seed(Bc * 7 + T + ND)
    whh = ((torch.rand(ND, 4 * H, H, device="cuda", generator=g) * 2 - 1) / np.sqrt(H) * 1.5).contiguous()
    G = (torch.randn(T * Bc, ND * 4 * H, device="cuda", generator=g) * 1.2).contiguous()
    packed = torch.empty(5 * ND * 4 * H * H, device="cuda", dtype=torch.float16)
    out = torch.full((T, Bc, ND * H), float("nan"), device="cuda")
    # the mixed form saves its gate activations as fp16, the fp32-parity form as fp32
    gates = torch.full((T * Bc, ND * 4 * H), float("nan"), device="cuda", dtype=torch.float32 if split else torch.float16)
    cs = torch.full((T, Bc, ND * H), float("nan"), device="cuda")
    N.check(N.lib().bci_selftest_rec_swap_fwd(_p(G), _p(whh), _p(packed), _p(out), _p(gates), _p(cs), Bc, T, ND, split, _stream()))
    torch.cuda.synchronize()
    want, wg, wc = _ref_forward64(G.double(), whh.double(), Bc, T, ND)
    assert torch.isfinite(out).all() and torch.isfinite(gates).all() and torch.isfinite(cs).all()
    e_h = float((out.double() - want).abs().max())
    e_g = float((gates.double().reshape(T, Bc, ND, H, 4) - wg).abs().max())
    e_c = float((cs.double().reshape(T, Bc, ND, H) - wc).abs().max())
    print(f"swap forward vs float64: h {e_h:.2e}, gates {e_g:.2e}, c {e_c:.2e} (Bc={Bc}, T={T}, ND={ND}, split={split})")
    if split:
        assert e_h <= 3e-6 and e_g <= 3e-6 and e_c <= 1e-5
    else:
        assert e_h <= 4e-3 and e_g <= 4e-3 and e_c <= 1.2e-2


@pytest.mark.parametrize("split", [0, 1])
@pytest.mark.parametrize("Bc,T,ND,decay", [(8, 6, 2, 0.0), (512, 16, 2, 0.0), (13, 30, 1, 0.0), (40, 200, 2, 0.08)])
def test_bptt_swap_matches_autograd(Bc, T, ND, decay, split):
    """split=0: bf16 operands (mixed mode); split=1: the fp32-parity form (dynamically scaled fp16 (hi, lo) pairs, three chains).
    decay > 0: the incoming gradient shrinks by e^(-decay t) along time (7 orders of magnitude over 200 steps), so the per-step
    scale of the split form has to follow the gradient's magnitude."""
    H = 128
    g = torch.Generator(device="cuda").manual_seed(Bc * 3 + T + ND)
    whh = ((torch.rand(ND, 4 * H, H, device="cuda", generator=g) * 2 - 1) / np.sqrt(H) * 1.5).contiguous()
    G = (torch.randn(T * Bc, ND * 4 * H, device="cuda", generator=g) * 1.2).contiguous()
    dout = (torch.randn(T, Bc, ND * H, device="cuda", generator=g) * 1e-3)
    if decay:
        dout = dout * torch.exp(-decay * torch.arange(T, device="cuda", dtype=torch.float32)).reshape(T, 1, 1)
    dout = dout.contiguous()
    G64 = G.double().requires_grad_(True)
    out, wg, wc = _ref_forward64(G64, whh.double(), Bc, T, ND)
    (out * dout.double()).sum().backward()
    want = G64.grad
    gates = wg.detach().to(torch.float32 if split else torch.float16).reshape(T * Bc, ND * 4 * H).contiguous()
    cs = wc.detach().float().reshape(T, Bc, ND * H).contiguous()
    packed = torch.empty(5 * ND * 4 * H * H, device="cuda", dtype=torch.float16)
    dG = torch.full((T * Bc, ND * 4 * H), float("nan"), device="cuda")
    N.check(N.lib().bci_selftest_bptt_swap(_p(dout), _p(gates), _p(cs), _p(whh), _p(packed), _p(dG), Bc, T, ND, split, _stream()))
    torch.cuda.synchronize()
    assert torch.isfinite(dG).all()
    # per time step, relative to that step's largest gradient (the steps differ by orders of magnitude when decay > 0)
    d3, w3 = dG.double().reshape(T, -1), want.reshape(T, -1)
    err = float(((d3 - w3).abs().max(dim=1).values / w3.abs().max(dim=1).values).max())
    cos = float((dG.double() * want).sum() / (dG.double().norm() * want.norm()))
    print(f"swap BPTT vs autograd: worst per-step rel-to-max err {err:.2e}, cosine {cos:.8f} (Bc={Bc}, T={T}, ND={ND}, decay={decay}, split={split})")
    if split:
        assert err <= 2e-5 and cos >= 0.9999999
    else:
        assert err <= 2e-2 and cos >= 0.9995


@pytest.mark.parametrize("Bc,T,ND", [(16, 8, 2), (512, 12, 2), (21, 40, 1), (100, 64, 2)])
def test_rec_swap256_forward_matches_float64(Bc, T, ND):
    """hidden_size 256: the CTA-pair form (each CTA owns 128 units, h halves pushed through DSMEM every step)."""
    H = 256
    g = torch.Generator(device="cuda").manual_seed(Bc * 7 + T + ND)
    whh = ((torch.rand(ND, 4 * H, H, device="cuda", generator=g) * 2 - 1) / np.sqrt(H) * 1.5).contiguous()
    G = (torch.randn(T * Bc, ND * 4 * H, device="cuda", generator=g) * 1.2).contiguous()
    packed = torch.empty(5 * ND * 4 * H * H, device="cuda", dtype=torch.float16)
    out = torch.full((T, Bc, ND * H), float("nan"), device="cuda")
    gates = torch.full((T * Bc, ND * 4 * H), float("nan"), device="cuda", dtype=torch.float16)
    cs = torch.full((T, Bc, ND * H), float("nan"), device="cuda")
    N.check(N.lib().bci_selftest_rec_swap256_fwd(_p(G), _p(whh), _p(packed), _p(out), _p(gates), _p(cs), Bc, T, ND, _stream()))
    torch.cuda.synchronize()
    want, wg, wc = _ref_forward64(G.double(), whh.double(), Bc, T, ND, H)
    assert torch.isfinite(out).all() and torch.isfinite(gates).all() and torch.isfinite(cs).all()
    e_h = float((out.double() - want).abs().max())
    e_g = float((gates.double().reshape(T, Bc, ND, H, 4) - wg).abs().max())
    e_c = float((cs.double().reshape(T, Bc, ND, H) - wc).abs().max())
    print(f"swap256 forward vs float64: h {e_h:.2e}, gates {e_g:.2e}, c {e_c:.2e} (Bc={Bc}, T={T}, ND={ND})")
    assert e_h <= 4e-3 and e_g <= 4e-3 and e_c <= 1.2e-2


@pytest.mark.parametrize("Bc,T,ND", [(16, 6, 2), (512, 10, 2), (21, 30, 1)])
def test_bptt_swap256_matches_autograd(Bc, T, ND):
    H = 256
    g = torch.Generator(device="cuda").manual_seed(Bc * 3 + T + ND)
    whh = ((torch.rand(ND, 4 * H, H, device="cuda", generator=g) * 2 - 1) / np.sqrt(H) * 1.5).contiguous()
    G = (torch.randn(T * Bc, ND * 4 * H, device="cuda", generator=g) * 1.2).contiguous()
    dout = (torch.randn(T, Bc, ND * H, device="cuda", generator=g) * 1e-3).contiguous()
    G64 = G.double().requires_grad_(True)
    out, wg, wc = _ref_forward64(G64, whh.double(), Bc, T, ND, H)
    (out * dout.double()).sum().backward()
    want = G64.grad
    gates = wg.detach().to(torch.float16).reshape(T * Bc, ND * 4 * H).contiguous()
    cs = wc.detach().float().reshape(T, Bc, ND * H).contiguous()
    packed = torch.empty(5 * ND * 4 * H * H, device="cuda", dtype=torch.float16)
    dG = torch.full((T * Bc, ND * 4 * H), float("nan"), device="cuda")
    N.check(N.lib().bci_selftest_bptt_swap256(_p(dout), _p(gates), _p(cs), _p(whh), _p(packed), _p(dG), Bc, T, ND, _stream()))
    torch.cuda.synchronize()
    assert torch.isfinite(dG).all()
    d3, w3 = dG.double().reshape(T, -1), want.reshape(T, -1)
    err = float(((d3 - w3).abs().max(dim=1).values / w3.abs().max(dim=1).values).max())
    cos = float((dG.double() * want).sum() / (dG.double().norm() * want.norm()))
    print(f"swap256 BPTT vs autograd: worst per-step rel-to-max err {err:.2e}, cosine {cos:.8f} (Bc={Bc}, T={T}, ND={ND})")
    assert err <= 2e-2 and cos >= 0.9995


@pytest.mark.parametrize("H,B,T", [(128, 16, 64), (128, 512, 256), (256, 24, 48), (256, 512, 64)])
def test_mixed_step_gradients_against_fp32_reference(H, B, T):
    """One training step in the mixed mode against torch autograd on the CPU port of the reference module (fp32)."""
    params = synth.make_lstm_params(46, 61, H, 3, logit_gain=4.0)
    x = synth.make_windows(12, B, T, 61)
    y = (np.arange(B) % 2).astype(np.int64)
    cw = np.array([0.8, 1.2], dtype=np.float32)
    port = torch_port.build_port(params, dropout=0.0)
    loss_ref, g_ref, _dx, _lg = torch_port.loss_and_grads(port, x, y, cw)
    m = lstm.from_params(params, precision="fp32", dropout=0.0).train()
    tr = train.FusedTrainer(m, lr=3e-4, weight_decay=1e-4, max_norm=1.0, class_weight=cw, precision="mixed")
    loss, norm = tr.step(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda())
    assert abs(float(loss) - loss_ref) <= 2e-3, (float(loss), loss_ref)
    worst_cos, worst_rel = 1.0, 0.0
    for k, _ in m.named_parameters():
        got, want = tr.grad_views[k].cpu().numpy().astype(np.float64).ravel(), g_ref[k].astype(np.float64).ravel()
        if k == "attention.attention.2.bias":     # softmax is shift-invariant: this gradient is exactly zero
            continue
        cos = float(got @ want / (np.linalg.norm(got) * np.linalg.norm(want) + 1e-300))
        rel = float(np.abs(got - want).max() / (np.abs(want).max() + 1e-300))
        worst_cos, worst_rel = min(worst_cos, cos), max(worst_rel, rel)
        assert cos >= 0.999 and rel <= 5e-2, (k, cos, rel)
    print(f"mixed step B={B} T={T}: loss {float(loss):.6f} vs {loss_ref:.6f}; worst cosine {worst_cos:.6f}, worst rel-to-max {worst_rel:.2e}")


def test_mixed_mode_through_autocast_autograd_bridge():
    """train_precision='auto': the reference's own loop under torch.autocast (04:486-490) takes the mixed step, without autocast the
    fp32-parity step; both produce gradients for every parameter."""
    H, B, T = 128, 8, 32
    params = synth.make_lstm_params(45, 61, H, 3, logit_gain=4.0)
    x = torch.from_numpy(synth.make_windows(10, B, T, 61)).cuda()
    y = (torch.arange(B) % 2).cuda()
    m = lstm.from_params(params, precision="auto", dropout=0.0).train()
    grads = {}
    for mode in ("fp32", "mixed"):
        m.zero_grad()
        if mode == "mixed":
            with torch.autocast("cuda", dtype=torch.float16):
                assert m._train_precision_now() == "mixed"
                loss = torch.nn.functional.cross_entropy(m(x).float(), y)
        else:
            assert m._train_precision_now() == "fp32"
            loss = torch.nn.functional.cross_entropy(m(x), y)
        loss.backward()
        grads[mode] = {k: p.grad.clone() for k, p in m.named_parameters()}
    diff = 0.0
    for k in grads["fp32"]:
        a, b = grads["fp32"][k].double().ravel(), grads["mixed"][k].double().ravel()
        if k == "attention.attention.2.bias":
            continue
        cos = float(a @ b / (a.norm() * b.norm() + 1e-300))
        assert cos >= 0.999, (k, cos)
        diff = max(diff, float((a - b).abs().max()))
    assert diff > 0.0     # the two modes really are different code paths


@pytest.mark.parametrize("H,B,T", [(128, 24, 48), (256, 24, 48)])
def test_mixed_step_with_dropout_matches_fp32_step(H, B, T):
    """dropout 0.4, same seed => same masks in both modes (the mask is a hash of seed / site / element index).  The tensor-core
    recurrences apply the inter-layer masks themselves (forward: dropped copy written next to h_t; BPTT: mask on the incoming
    gradient); at H = 256 the fp32-parity step still uses the separate mask passes of round 1, so this also cross-checks the two
    implementations of the same dropout."""
    params = synth.make_lstm_params(47, 61, H, 3, logit_gain=4.0)
    x = torch.from_numpy(synth.make_windows(13, B, T, 61)).cuda()
    y = (torch.arange(B) % 2).cuda()
    grads, losses = {}, {}
    for mode in ("fp32", "mixed"):
        m = lstm.from_params(params, precision="fp32", dropout=0.4).train()
        tr = train.FusedTrainer(m, lr=0.0, weight_decay=0.0, max_norm=0.0, precision=mode)
        loss, _ = tr.step(x, y, seed=99)
        losses[mode] = float(loss)
        grads[mode] = {k: v.clone() for k, v in tr.grad_views.items()}
        tr.close()
    assert abs(losses["fp32"] - losses["mixed"]) <= 2e-3, losses
    for k in grads["fp32"]:
        if k == "attention.attention.2.bias":
            continue
        a, b = grads["fp32"][k].double().ravel(), grads["mixed"][k].double().ravel()
        cos = float(a @ b / (a.norm() * b.norm() + 1e-300))
        assert cos >= 0.999, (k, cos)


@pytest.mark.parametrize("kw", [dict(hidden_size=128, bidirectional=False, use_attention=False, num_layers=1),
                                dict(hidden_size=256, bidirectional=True, use_attention=False, use_layer_norm=False, num_layers=2),
                                dict(hidden_size=256, bidirectional=False, use_attention=True, num_layers=3)])
def test_mixed_mode_on_ablation_variants(kw):
    """The ablation switches of 09_sensitivity_analysis.py:176-240 (unidirectional, mean pooling, no LayerNorm, fewer layers) under the
    mixed training mode: gradients against the fp32-parity step of the same variant."""
    torch.manual_seed(3)
    m = lstm.AblationLSTMModel(input_size=61, dropout=0.0, **kw).cuda().train()
    B, T = 20, 40
    x = torch.from_numpy(synth.make_windows(14, B, T, 61)).cuda()
    y = (torch.arange(B) % 2).cuda()
    grads = {}
    for mode in ("fp32", "mixed"):
        m.train_precision = mode
        m.zero_grad()
        torch.nn.functional.cross_entropy(m(x), y).backward()
        grads[mode] = {k: p.grad.clone() for k, p in m.named_parameters()}
    moved = 0.0
    for k in grads["fp32"]:
        a, b = grads["fp32"][k].double().ravel(), grads["mixed"][k].double().ravel()
        if float(a.norm()) < 1e-12 or k == "attention.attention.2.bias":   # softmax is shift-invariant: that gradient is rounding noise
            continue
        cos = float(a @ b / (a.norm() * b.norm() + 1e-300))
        assert cos >= 0.999, (k, cos)
        moved = max(moved, float((a - b).abs().max()))
    assert moved > 0.0


_JITTER_CODE = """
import numpy as np, torch
from lstm_ode_bci_b200 import lstm, synth, train
out = []
for H, mode in ((128, "fp32"), (128, "mixed"), (256, "mixed")):
    params = synth.make_lstm_params(48, 61, H, 3, logit_gain=4.0)
    x = torch.from_numpy(synth.make_windows(15, 40, 96, 61)).cuda()
    y = (torch.arange(40) % 2).cuda()
    m = lstm.from_params(params, precision="fp32", dropout=0.3).train()
    tr = train.FusedTrainer(m, lr=0.0, weight_decay=0.0, max_norm=0.0, precision=mode)
    loss, _ = tr.step(x, y, seed=5)
    g = tr.grad.double()
    out += [float(loss), float(g.norm()), float(g.abs().max())]
    tr.close()
m32 = lstm.from_params(synth.make_lstm_params(48, 61, 128, 3, logit_gain=4.0), precision="bf16")
out.append(float(m32.predict_proba(torch.from_numpy(synth.make_windows(16, 100, 64, 61)).cuda()).double().sum()))
print("RESULT", " ".join("%.12e" % v for v in out))
"""


def test_swapped_recurrences_are_robust_to_timing_jitter():
    """BCI_FUSED_JITTER on the swapped recurrences: every MMA warp and epilogue warp sleeps a pseudo-random time (up to 4 us) before
    its waits, arrives and DSMEM pushes -- single-CTA kernels (two mbarriers) and the H = 256 CTA pairs (double-buffered B tiles and
    x_in barriers, re-armed by the waiter).  Results must be those of the unperturbed run (up to the summation order of the
    bias-gradient atomics)."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    vals = {}
    for tag, extra in (("plain", {}), ("jitter", {"BCI_FUSED_JITTER": "4096"})):
        env = dict(os.environ, PYTHONPATH=root, **extra)
        r = subprocess.run([sys.executable, "-c", _JITTER_CODE], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "RESULT" in r.stdout, r.stderr[-2000:]
        vals[tag] = np.array([float(v) for v in r.stdout.split("RESULT")[1].split()])
    assert np.isfinite(vals["jitter"]).all()
    rel = np.abs(vals["jitter"] - vals["plain"]) / (np.abs(vals["plain"]) + 1e-30)
    assert rel.max() <= 1e-5, (rel, vals)
