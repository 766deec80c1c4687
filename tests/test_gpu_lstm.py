"""GPU parity of the BiLSTM + attention-pooling forward (through the C ABI / custom op).
fp32 mode tolerance (north_star): max-abs <= 1e-5 on logits and P(open)/P(closed); attention <= 1e-6.
bf16 (tcgen05) mode tolerance, stated separately: logits <= 3e-2 * gain-scale, probabilities <= 1e-2,
attention <= 2e-3 (calibrated against the reference's own autocast-vs-fp32 gap)."""
import numpy as np
import pytest
import torch

from lstm_ode_bci_b200 import lstm, synth
from oracle import lstm_oracle, torch_port

pytestmark = pytest.mark.gpu


def _inputs(g):
    if "x" in g:
        return {k[6:]: v for k, v in g.items() if k.startswith("param:")}, g["x"]
    params = synth.make_lstm_params(int(g["seed_w"]), int(g["C"]), int(g["H"]), int(g["L"]), logit_gain=float(g["gain"]))
    x = synth.make_windows(int(g["seed_x"]), int(g["B"]), int(g["T"]), int(g["C"]), structured=bool(g["structured"]))
    return params, x


@pytest.mark.parametrize("name", ["lstm_h128.npz", "lstm_h128_t64.npz", "lstm_h256.npz"])
def test_fp32_forward_matches_reference_golden(golden, name):
    g = golden(name)
    params, x = _inputs(g)
    m = lstm.from_params(params, precision="fp32")
    with torch.no_grad():
        logits, attn = m(torch.from_numpy(x).cuda(), return_attention=True)
        probs = m.predict_proba(torch.from_numpy(x).cuda())
    assert np.abs(logits.cpu().numpy() - g["logits"]).max() <= 1e-5
    assert np.abs(probs.cpu().numpy() - g["probs"]).max() <= 1e-5
    assert np.abs(attn.cpu().numpy() - g["attention"]).max() <= 1e-6
    assert np.abs(attn.sum(dim=1).cpu().numpy() - 1).max() <= 1e-5


@pytest.mark.parametrize("H,B,T", [(128, 32, 256), (128, 1, 256), (128, 33, 40), (256, 17, 64), (128, 70, 7)])
def test_fp32_forward_matches_oracle(H, B, T):
    """config 1 (B=32 x 256 x 61) plus ragged tiles, odd sequence lengths and H=256."""
    params = synth.make_lstm_params(100 + H + B, 61, H, 3, logit_gain=6.0)
    x = synth.make_windows(200 + B, B, T, 61, structured=True)
    port = torch_port.build_port(params).eval()
    with torch.no_grad():
        want_logits, want_attn = port(torch.from_numpy(x), return_attention=True)
        want_probs = torch.softmax(want_logits, dim=1)
    m = lstm.from_params(params, precision="fp32")
    with torch.no_grad():
        logits, attn = m(torch.from_numpy(x).cuda(), return_attention=True)
        probs = m.predict_proba(torch.from_numpy(x).cuda())
    assert np.abs(logits.cpu().numpy() - want_logits.numpy()).max() <= 1e-5
    assert np.abs(probs.cpu().numpy() - want_probs.numpy()).max() <= 1e-5
    assert np.abs(attn.cpu().numpy() - want_attn.numpy()).max() <= 1e-6
    if B <= 2:  # independent float64 numpy restatement as a second witness
        lg64, at64 = lstm_oracle.forward(params, x)
        assert np.abs(logits.cpu().numpy() - lg64).max() <= 1e-5
        assert np.abs(attn.cpu().numpy() - at64).max() <= 1e-6


def test_fp32_chunked_batch_equals_unchunked():
    """a batch == the concatenation of smaller calls; windows are independent (SURVEY.md §8 e).  Within one recurrence kernel a
    window's result is BIT-identical whatever batch it arrives in: the tensor-core pair recurrence (>= 16 work items of 256
    windows x direction, lstm_fp32_tc.cu) and the CUDA-core recurrence (smaller batches); across the two the results agree to
    fp32 rounding (both are <= 1e-5 from the reference)."""
    params = synth.make_lstm_params(5, 61, 128, 3, logit_gain=6.0)
    m = lstm.from_params(params, precision="fp32")
    x = torch.from_numpy(synth.make_windows(6, 4200, 16, 61)).cuda()
    with torch.no_grad():
        full = m(x)
        parts = torch.cat([m(x[:2100]), m(x[2100:])])           # tensor-core recurrence in all three calls
        small = torch.cat([m(x[:1000]), m(x[1000:1500])])       # CUDA-core recurrence
        small2 = torch.cat([m(x[:700]), m(x[700:1500])])
    assert torch.equal(full, parts)
    assert torch.equal(small, small2)
    assert float((full[:1500] - small).abs().max()) <= 2e-6
    assert m(x[:0]).shape == (0, 2)


def test_fp32_tensorcore_recurrence_matches_oracle():
    """fp32 parity mode at a batch that runs the recurrence on the tensor cores (lstm_fp32_tc.cu: split fp16 MMAs on CTA pairs;
    2304 windows = 9 pairs x 2 directions = 18 work items), full sequence length (3 x 256 dependent steps), logit gain 12: the
    north_star tolerances against the ORACLE (torch CPU port of the reference module) on three 32-window slices, one of them in
    the ragged last pair."""
    B, T = 2304 + 40, 256
    params = synth.make_lstm_params(42, 61, 128, 3, logit_gain=12.0)
    x = synth.make_windows(21, B, T, 61, structured=True)
    m = lstm.from_params(params, precision="fp32")
    xc = torch.from_numpy(x).cuda()
    with torch.no_grad():
        logits, attn = m(xc, return_attention=True)
        probs = m.predict_proba(xc)
    port = torch_port.build_port(params).eval()
    for lo in (0, 1100, B - 32):
        with torch.no_grad():
            wl, wa = port(torch.from_numpy(x[lo:lo + 32]), return_attention=True)
            wp = torch.softmax(wl, 1)
        dl = float((logits[lo:lo + 32].cpu() - wl).abs().max())
        dp = float((probs[lo:lo + 32].cpu() - wp).abs().max())
        da = float((attn[lo:lo + 32].cpu() - wa).abs().max())
        print(f"fp32 tensor-core recurrence vs oracle, windows {lo}..{lo + 32}: dlogit {dl:.2e} dprob {dp:.2e} dattn {da:.2e}")
        assert dl <= 1e-5 and dp <= 1e-5 and da <= 1e-6, (lo, dl, dp, da)


def test_fp32_full_pass_matches_oracle_and_is_order_independent():
    """The fp32 parity mode at the bench's own pass size (`bci_lstm_chunk_windows`: 9 472 windows on a 148-SM B200 = 148 work items on 74
    CTA pairs, two per pair -- the staggered recurrence lstm_rec_f16x3_pipe with its per-item reload of the weights and barrier
    parities carried across items): windows are independent, so reversing their order reverses the outputs BIT FOR BIT; three
    32-window slices (first pair, a second-round item, the last pair) against the ORACLE within the north_star tolerances."""
    from lstm_ode_bci_b200 import ops
    params = synth.make_lstm_params(42, 61, 128, 3, logit_gain=12.0)
    m = lstm.from_params(params, precision="fp32")
    B = ops.lstm_chunk_windows(m._engine("fp32"))
    g = torch.Generator(device="cuda").manual_seed(17)
    x = torch.randn((B, 256, 61), device="cuda", generator=g)
    with torch.no_grad():
        p, a = m.predict_proba(x, return_attention=True)
        pr, ar = m.predict_proba(x.flip(0).contiguous(), return_attention=True)
    assert torch.isfinite(p).all() and torch.isfinite(a).all()
    assert torch.equal(pr.flip(0), p) and torch.equal(ar.flip(0), a)
    port = torch_port.build_port(params).eval()
    xc = x.cpu()
    for lo in (0, (3 * B) // 4 - 16, B - 32):
        with torch.no_grad():
            wl, wa = port(xc[lo:lo + 32], return_attention=True)
            wp = torch.softmax(wl, 1)
        dp = float((p[lo:lo + 32].cpu() - wp).abs().max())
        da = float((a[lo:lo + 32].cpu() - wa).abs().max())
        print(f"fp32 full pass vs oracle, windows {lo}..{lo + 32}: dprob {dp:.2e} dattn {da:.2e}")
        assert dp <= 1e-5 and da <= 1e-6, (lo, dp, da)


def test_weights_reload_after_update():
    params = synth.make_lstm_params(9, 61, 128, 3)
    m = lstm.from_params(params, precision="fp32")
    x = torch.from_numpy(synth.make_windows(3, 4, 32, 61)).cuda()
    with torch.no_grad():
        a = m(x).clone()
        m.classifier[6].bias.add_(1.0)          # in-place update bumps the version -> repack
        b = m(x)
    assert np.allclose((b - a).cpu().numpy(), 1.0, atol=1e-6)


def test_state_dict_roundtrip_with_port():
    """A checkpoint of the reference layout loads unchanged (SURVEY.md §5 checkpoint row)."""
    params = synth.make_lstm_params(11, 61, 128, 3)
    port = torch_port.build_port(params)
    m = lstm.EnhancedLSTMModel(61, 128, 3, 2).cuda().eval()
    m.load_state_dict(port.state_dict(), strict=True)
    x = synth.make_windows(4, 3, 48, 61)
    with torch.no_grad():
        want = port.eval()(torch.from_numpy(x))
        got = m(torch.from_numpy(x).cuda())
    assert np.abs(got.cpu().numpy() - want.numpy()).max() <= 1e-5


def test_environment_switches_select_equivalent_paths(tmp_path):
    """BCI_FP32_GEMM=simt (CUDA-core GEMMs) and BCI_BF16_POOL=two (two-kernel pooling) are read once per process: each is run
    in a subprocess on the same seeded inputs and compared with this process's default path."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    prog = (
        "import sys, numpy as np, torch\n"
        "sys.path.insert(0, %r)\n"
        "from lstm_ode_bci_b200 import lstm, synth\n"
        "prec, B, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]\n"
        "m = lstm.from_params(synth.make_lstm_params(42, 61, 128, 3, logit_gain=12.0), precision=prec)\n"
        "x = torch.from_numpy(synth.make_windows(5, B, 256, 61, structured=True)).cuda()\n"
        "with torch.no_grad():\n"
        "    p = m.predict_proba(x)\n"
        "np.save(out, p.cpu().numpy())\n" % root)
    script = tmp_path / "run.py"
    script.write_text(prog)

    def run(prec, B, env):
        out = str(tmp_path / ("%s_%d_%s.npy" % (prec, B, "_".join(env) or "default")))
        e = dict(os.environ)
        e.update(env)
        subprocess.run([sys.executable, str(script), prec, str(B), out], check=True, env=e, timeout=300)
        return np.load(out)

    a, b = run("fp32", 8, {}), run("fp32", 8, {"BCI_FP32_GEMM": "simt"})      # 2048 rows: the tcgen05 GEMMs are used by default
    assert np.abs(a - b).max() <= 1e-5
    c, d = run("bf16", 4224, {}), run("bf16", 4224, {"BCI_BF16_POOL": "two"})
    assert np.abs(c - d).max() <= 2e-3
