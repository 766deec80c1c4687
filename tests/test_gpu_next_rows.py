"""SURVEY.md §8 f rows built so far: (1) attribution gradients, (2) ODE parameter fitting."""
import numpy as np
import pytest
import torch

from lstm_ode_bci_b200 import explain, lstm, ode, synth
from oracle import ode_oracle as oo
from oracle import torch_port

pytestmark = pytest.mark.gpu


def test_batched_input_gradients_equal_per_sample_backward_of_reference_port():
    """07:242-258: outputs[i, pred_i].backward(retain_graph=True) per sample == row i of one batched backward."""
    params = synth.make_lstm_params(21, 61, 128, 3, logit_gain=6.0)
    x = synth.make_windows(22, 5, 24, 61)
    port = torch_port.build_port(params, dropout=0.0).train()
    xt = torch.from_numpy(x).requires_grad_(True)
    out = port(xt)
    pred = out.argmax(dim=1)
    want = np.empty_like(x)
    for i in range(len(x)):
        if xt.grad is not None:
            xt.grad.zero_()
        out[i, pred[i]].backward(retain_graph=True)
        want[i] = xt.grad[i].numpy()
    m = lstm.from_params(params, precision="fp32", dropout=0.0).train()
    got = explain.input_gradients(m, x, batch_size=3)
    assert np.abs(got - want).max() <= 2e-4 * np.abs(want).max()
    # the reference's own per-sample pattern against the drop-in module (repeated backward on one forward)
    xc = torch.from_numpy(x).cuda().requires_grad_(True)
    outc = m(xc)
    for i in (0, 3):
        if xc.grad is not None:
            xc.grad.zero_()
        outc[i, pred[i]].backward(retain_graph=True)
        assert np.abs(xc.grad[i].cpu().numpy() - want[i]).max() <= 2e-4 * np.abs(want).max()
    np.random.seed(0)
    df = explain.compute_channel_importance(m, x, n_samples=4, batch_size=2)
    assert abs(df["Importance"].sum() - 1) < 1e-6 and len(df) == 61


def test_population_loss_and_fit_recover_known_rates():
    true = {"k_ap": 0.12, "k_af": 0.03, "k_pa": 0.2, "k_pf": 0.05, "k_fa": 0.08, "k_fp": 0.15}
    tp = np.linspace(0, 40, 41)
    k = oo.rates_to_array(true)[:, None]
    obs = oo.exact_solution(oo.STYLE_REF06, [[0.7, 0.2, 0.1]], k, 40.0, 41)[0]
    model = ode.CognitiveStateODE()
    # population objective == the reference's scalar objective (05:259-283) evaluated per candidate
    rng = np.random.default_rng(0)
    lo = np.array([b[0] for b in ode.CognitiveStateODE.FIT_BOUNDS]); hi = np.array([b[1] for b in ode.CognitiveStateODE.FIT_BOUNDS])
    pop = rng.uniform(lo[:, None], hi[:, None], size=(6, 33))
    got = model.population_loss(pop, obs, tp)
    sol = oo.exact_solution(oo.STYLE_REF06, np.repeat(obs[:1], 33, axis=0), pop, 40.0, 41)
    want = ((sol - obs[None]) ** 2).mean(axis=(1, 2)) + 0.001 * (pop ** 2).sum(axis=0)
    assert np.abs(got - want).max() <= 1e-6
    fitted, loss = model.fit_to_data(obs, tp)
    floor = 0.001 * float((k ** 2).sum())                     # regulariser at the true rates; data term ~0 there
    assert loss <= floor + 1e-6
    t, sol = ode.CognitiveStateODE(fitted).solve(obs[0], (0, 40), 41)
    assert np.abs(sol - obs).max() <= 2e-2                    # the 1e-3 |k|^2 regulariser trades a little data fit for smaller rates


def test_fit_matches_reference_seeded_fit(golden):
    """The reference's own seeded fit (05_ode_model.py:296-303: differential_evolution(seed=42, maxiter=1000, tol=1e-7,
    polish=True)), run by the live reference (tests/golden/make_golden_fit.py): the GPU objective reproduces the reference's
    final loss at the reference's fitted rates, and the GPU fit converges to the same loss within 1e-6."""
    g = golden("ode_ref05_fit.npz")
    obs, tp = g["observed"], g["time_points"]
    model = ode.CognitiveStateODE()
    at_ref = model.population_loss(g["fitted"].reshape(6, 1), obs, tp)[0]
    assert abs(at_ref - float(g["loss"])) <= 1e-7, (at_ref, float(g["loss"]))
    fitted, loss = model.fit_to_data(obs, tp)
    print("GPU fit loss %.10f, reference %.10f" % (loss, float(g["loss"])))
    assert abs(loss - float(g["loss"])) <= 1e-6
    assert all(lo - 1e-12 <= fitted[k] <= hi + 1e-12 for k, (lo, hi) in zip(ode.RATE_ORDER, ode.CognitiveStateODE.FIT_BOUNDS))
