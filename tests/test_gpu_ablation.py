"""GPU parity of the ablation variants of the LSTM op (SURVEY.md §8 f row 3; 09_sensitivity_analysis.py:176-240, 330-378):
unidirectional, 1-2 layers, mean pooling instead of attention, LayerNorm off, H = 128 and 256 -- forward and the training
step, against the golden outputs of the live reference's AblationLSTMModel and torch autograd on the CPU port.
Tolerances: logits <= 1e-5, loss <= 1e-5, gradients <= 2e-4 of each tensor's max-abs (as in test_gpu_train.py)."""
import numpy as np
import pytest
import torch

from lstm_ode_bci_b200 import lstm
from lstm_ode_bci_b200._native import BciError
from oracle import lstm_oracle, torch_port
from test_oracle_golden import ABLATION_TAGS, ablation_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ABLATION_TAGS)
def test_forward_and_gradients_match_reference(golden, tag):
    g = golden("ablation_ref09.npz")
    params, x, y = ablation_case(g, tag)
    # eval forward (inference path)
    m = lstm.from_params(params, precision="fp32", dropout=0.0)
    with torch.no_grad():
        logits, attn = m(torch.from_numpy(x).cuda(), return_attention=True)
    assert np.abs(logits.cpu().numpy() - g[tag + ":logits"]).max() <= 1e-5
    _, attn_ref = lstm_oracle.forward(params, x)
    assert np.abs(attn.cpu().numpy() - attn_ref).max() <= 1e-6
    # training step (autograd bridge): reference golden summaries + full gradients of the CPU port
    m.train()
    xc = torch.from_numpy(x).cuda().requires_grad_(True)
    out = m(xc)
    loss = torch.nn.functional.cross_entropy(out, torch.from_numpy(y).cuda())
    loss.backward()
    assert np.abs(out.detach().cpu().numpy() - g[tag + ":logits"]).max() <= 1e-5
    assert abs(float(loss.detach()) - float(g[tag + ":loss"])) <= 1e-5
    port = torch_port.build_port(params, dropout=0.0)
    _, g_ref, dx_ref, _ = torch_port.loss_and_grads(port, x, y)
    for k, p in m.named_parameters():
        got, want = p.grad.cpu().numpy(), g_ref[k]
        tol = 2e-4 * np.abs(want).max() + 1e-7
        assert np.abs(got - want).max() <= tol, (k, np.abs(got - want).max(), tol)
        ref_norm = float(g[tag + ":gnorm:" + k])
        assert abs(np.linalg.norm(got.astype(np.float64)) - ref_norm) <= 3e-4 * max(ref_norm, 1e-4), k
    assert np.abs(xc.grad.cpu().numpy() - dx_ref).max() <= 2e-4 * np.abs(dx_ref).max() + 1e-8


def test_ablation_model_class_mirrors_09():
    """Constructor defaults and forward contract of 09's AblationLSTMModel; state dict keys of each variant."""
    m = lstm.AblationLSTMModel(input_size=61, hidden_size=256, num_layers=1, bidirectional=False, use_attention=False).cuda().eval()
    keys = set(m.state_dict().keys())
    assert "lstm.weight_ih_l0" in keys and "lstm.weight_ih_l0_reverse" not in keys
    assert not any(k.startswith("attention") for k in keys) and "layer_norm.weight" in keys
    x = torch.randn(3, 32, 61, device="cuda")
    with torch.no_grad():
        out = m(x)
    assert tuple(out.shape) == (3, 2)
    port = torch_port.build_port({k: v.cpu().numpy() for k, v in m.state_dict().items()}, dropout=0.0).eval()
    with torch.no_grad():
        want = port(x.cpu())
    assert (out.cpu() - want).abs().max() <= 1e-5
    m2 = lstm.AblationLSTMModel(use_layer_norm=False)
    assert "layer_norm.weight" not in m2.state_dict() and "input_proj.1.weight" not in m2.state_dict()


def test_variants_are_fp32_only():
    m = lstm.AblationLSTMModel(hidden_size=128, bidirectional=False, precision="bf16").cuda().eval()
    with pytest.raises(BciError):
        m(torch.randn(2, 128, 61, device="cuda"))
