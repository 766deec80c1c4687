"""SURVEY.md §8 f rows built so far: (1) attribution gradients and permutation importance (07), (2) ODE parameter fitting and
sensitivity analysis (05)."""
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


@pytest.mark.parametrize("n,T,C", [(7, 256, 61), (5, 6, 3), (3, 9, 4)])
def test_permute_channels_kernel_is_the_reference_gather(n, T, C):
    """bci_permute_channels against the host-side copy of 07:336-339 (oracle/explain_oracle.permuted_copy): bit-exact in fp32,
    and in bf16 bit-exact with torch's round-to-nearest-even narrowing; vector (T*C % 4 == 0) and scalar paths, an
    unpermuted variant (channel < 0), row ranges that start and end inside a variant, an empty range."""
    from lstm_ode_bci_b200 import ops
    from oracle import explain_oracle
    rng = np.random.default_rng(n * 100 + C)
    x = rng.standard_normal((n, T, C)).astype(np.float32)
    channels = np.array([C - 1, -1, 0, C // 2, 1 % C], dtype=np.int32)
    perms = np.stack([rng.permutation(n) for _ in channels]).astype(np.int32)
    want = np.concatenate([x if ch < 0 else explain_oracle.permuted_copy(x, p, ch) for ch, p in zip(channels, perms)])
    xd = torch.from_numpy(x).cuda()
    pd_, cd = torch.from_numpy(perms).cuda().view(-1), torch.from_numpy(channels).cuda()
    total = len(channels) * n
    got = ops.permute_channels(xd, pd_, cd, 0, total)
    assert np.array_equal(got.cpu().numpy(), want)
    r0, rows = n // 2 + 1, 3 * n                                       # starts and ends inside a variant
    part = ops.permute_channels(xd, pd_, cd, r0, rows)
    assert np.array_equal(part.cpu().numpy(), want[r0:r0 + rows])
    b16 = ops.permute_channels(xd, pd_, cd, r0, rows, bf16_out=True)
    assert b16.dtype == torch.bfloat16
    assert torch.equal(b16.cpu(), torch.from_numpy(want[r0:r0 + rows]).to(torch.bfloat16))
    assert ops.permute_channels(xd, pd_, cd, total, 0).shape == (0, T, C)
    with pytest.raises(Exception):
        ops.permute_channels(xd, pd_, cd, total - 1, 2)


def test_permutation_importance_matches_reference_07(golden):
    """07_explainability.py:287-361 run by the LIVE reference on a seeded model / test set (tests/golden/make_golden_explain.py)
    against explain.compute_permutation_importance seeded the same way.  The subset and the permutations are identical by
    construction; a prediction can differ only where a permuted window's logit margin is inside the fp32 mode's 1e-5 (x gain 200)
    of zero, and each such flip moves one channel's importance by 1 / (n_samples * n_permutations): at most three allowed."""
    from test_oracle_golden import explain_case_inputs
    g = golden("explain_ref07.npz")
    params, X, y = explain_case_inputs(g)
    n, reps = int(g["n_samples"]), int(g["n_permutations"])
    m = lstm.from_params(params, precision="fp32")
    assert np.abs(m(torch.from_numpy(X).cuda()).cpu().numpy() - g["logits"]).max() <= 2e-3       # logits of O(10) at gain 200
    np.random.seed(int(g["numpy_seed"]))
    df = explain.compute_permutation_importance(m, X, y, n_permutations=reps, n_samples=n, batch_size=32)
    assert list(df.columns) == ["Channel", "Importance"] and len(df) == 61
    assert df["Importance"].is_monotonic_decreasing
    imp = np.empty(61)
    imp[[int(c[2:]) - 1 for c in df["Channel"]]] = df["Importance"].to_numpy()
    flips = np.abs(imp - g["importance"]).sum() * n * reps
    print("flipped predictions vs the reference run: %.1f" % flips)
    assert flips <= 3.0 + 1e-9
    assert set(np.argsort(-imp)[:3]) == set(g["sorted_order"][:3])
    # small gathers (several launches, variants split across passes) give the same counts
    np.random.seed(int(g["numpy_seed"]))
    idx = np.random.choice(len(X), n, replace=False)
    perms = np.stack([np.random.permutation(n) for _ in range(6)])
    chans = [3, 3, 17, -1, 40, 5]
    a = explain.permuted_channel_accuracy(m, X[idx], y[idx], chans, perms)
    b = explain.permuted_channel_accuracy(m, X[idx], y[idx], chans, perms, rows_per_pass=n + 7)
    assert np.array_equal(a, b)
    # bf16 engine (bf16 gather): the same channels dominate, importances within a few flips
    mb = lstm.from_params(params, precision="bf16")
    np.random.seed(int(g["numpy_seed"]))
    dfb = explain.compute_permutation_importance(mb, X, y, n_permutations=reps, n_samples=n)
    impb = np.empty(61)
    impb[[int(c[2:]) - 1 for c in dfb["Channel"]]] = dfb["Importance"].to_numpy()
    print("bf16 engine: max |importance - reference| = %.4f" % np.abs(impb - g["importance"]).max())
    assert set(np.argsort(-impb)[:2]) == set(g["sorted_order"][:2]) and g["sorted_order"][2] in np.argsort(-impb)[:5]
    assert np.abs(impb - g["importance"]).max() <= 0.1       # bf16 logits at gain 200 are ~3e-2 off: a few flips per channel
    # the patched reference labels channels with its own EEG_CHANNELS (07:222-225): channel_names passes through
    names = ["E%d" % i for i in range(61)]
    np.random.seed(1)
    dfn = explain.compute_permutation_importance(m, X[:8], y[:8], n_permutations=1, n_samples=8, channel_names=names)
    assert set(dfn["Channel"]) == set(names)


def test_sensitivity_analysis_matches_reference_05(golden):
    """05_ode_model.py:687-750 and get_steady_state (05:198-221) run by the live reference (LSODA to t = 1000) against the
    one-launch fp64 ensemble: steady states within 1e-7, sensitivities within 2e-5 (the reference's own integration error
    amplified by 1 / (0.4 k)); result records in the reference's form and order."""
    g = golden("ode_ref05_sensitivity.npz")
    for j in (0, 1):
        base = {k: float(g["params_%d" % j][i]) for i, k in enumerate(synth.RATE_ORDER)}
        model = ode.CognitiveStateODE(dict(base))
        res = ode.sensitivity_analysis(model, None)
        assert [r["parameter"] for r in res] == list(g["names_%d" % j])
        assert set(res[0]) == {"parameter", "sens_Active", "sens_Passive", "sens_Fatigued"}
        got = np.array([[r["sens_Active"], r["sens_Passive"], r["sens_Fatigued"]] for r in res])
        assert np.abs(got - g["sens_%d" % j]).max() <= 2e-5, np.abs(got - g["sens_%d" % j]).max()
        assert np.abs(got.sum(axis=1)).max() <= 1e-9                       # A + P + F is conserved: sensitivities sum to 0
        ss = ode.steady_state_ensemble(g["params_%d" % j].reshape(6, 1))[0]
        assert np.abs(ss - g["steady_%d" % j]).max() <= 1e-7
        one = model.get_steady_state()                                     # the object's own (fp32 RK4) path agrees to fp32 accuracy
        assert abs(one["Active"] - ss[0]) <= 2e-6 and abs(one["Fatigued"] - ss[2]) <= 2e-6
        assert model.params == base                                        # the caller's parameters are left untouched
