"""Pin the CPU oracle against outputs of the live reference (tests/golden/*.npz).

The reference has no tests of its own; these vectors were produced by
tests/golden/make_golden.py importing /root/reference in the build container.
"""
import numpy as np
import pytest

from lstm_ode_bci_b200 import synth
from oracle import lstm_oracle, ode_oracle, torch_port

LSTM_CASES = ["lstm_tiny.npz", "lstm_h128_t64.npz", "lstm_h128.npz", "lstm_h256.npz"]


def _lstm_inputs(g):
    if "x" in g:
        params = {k[6:]: v for k, v in g.items() if k.startswith("param:")}
        return params, g["x"]
    params = synth.make_lstm_params(int(g["seed_w"]), int(g["C"]), int(g["H"]), int(g["L"]),
                                    logit_gain=float(g["gain"]))
    x = synth.make_windows(int(g["seed_x"]), int(g["B"]), int(g["T"]), int(g["C"]),
                           structured=bool(g["structured"]))
    return params, x


@pytest.mark.parametrize("name", LSTM_CASES)
def test_numpy_oracle_matches_reference_forward(golden, name):
    g = golden(name)
    params, x = _lstm_inputs(g)
    if name == "lstm_h256.npz":
        x = x[:2]
    logits, attn = lstm_oracle.forward(params, x, dtype=np.float64)
    n = len(x)
    # reference ran fp32; float64 oracle differs only by the reference's own rounding
    assert np.abs(logits - g["logits"][:n]).max() < 2e-6
    assert np.abs(attn - g["attention"][:n]).max() < 2e-7
    assert np.abs(lstm_oracle.softmax_probs(logits) - g["probs"][:n]).max() < 1e-6
    assert np.allclose(attn.sum(axis=1), 1.0, atol=1e-12)


@pytest.mark.parametrize("name", ["lstm_tiny.npz", "lstm_h128_t64.npz"])
def test_torch_port_matches_reference_forward(golden, name):
    import torch
    g = golden(name)
    params, x = _lstm_inputs(g)
    m = torch_port.build_port(params).eval()
    with torch.no_grad():
        logits, attn = m(torch.from_numpy(x), return_attention=True)
    assert np.abs(logits.numpy() - g["logits"]).max() < 1e-6
    assert np.abs(attn.numpy() - g["attention"]).max() < 1e-7
    # state-dict key ABI
    assert list(m.state_dict().keys()) == list(synth.lstm_param_shapes(
        int(g["C"]), int(g["H"]), int(g["L"])).keys())


@pytest.mark.parametrize("name", ["lstm_grad_tiny.npz", "lstm_grad_h128.npz"])
def test_torch_port_matches_reference_gradients(golden, name):
    g = golden(name)
    params = synth.make_lstm_params(int(g["seed_w"]), int(g["C"]), int(g["H"]), int(g["L"]), logit_gain=float(g["gain"]))
    x = synth.make_windows(int(g["seed_x"]), int(g["B"]), int(g["T"]), int(g["C"]))
    m = torch_port.build_port(params, dropout=0.0)
    loss, grads, dx, _ = torch_port.loss_and_grads(m, x, g["y"], g["class_weight"])
    assert abs(loss - float(g["loss"])) < 1e-6
    for k, gr in grads.items():
        ref_norm = float(g["gnorm:" + k])
        assert abs(np.linalg.norm(gr.astype(np.float64)) - ref_norm) <= 1e-5 * max(ref_norm, 1e-3), k
        assert np.allclose(gr.reshape(-1)[:16], g["ghead:" + k], rtol=1e-4, atol=1e-7), k
    assert abs(np.linalg.norm(dx) - float(g["dx_norm"])) < 1e-5 * float(g["dx_norm"])


def test_coupling_and_initial_state_match_reference(golden):
    g = golden("ode_ref06.npz")
    k = ode_oracle.modulate_rates(g["base"], g["alpha"], g["p_closed"], g["p_open"])
    assert np.array_equal(k, g["rates"])          # bit-exact, incl. float32 promotion + floor
    assert np.array_equal(ode_oracle.initial_state_06(g["p_open"], g["p_closed"]), g["y0"])
    # alpha = 0 leaves the rates untouched (above the floor)
    k0 = ode_oracle.modulate_rates(g["base"], 0.0, g["p_closed"], g["p_open"])
    assert np.array_equal(k0, np.maximum(0.001, g["base"].astype(np.float32).astype(np.float64)))


def test_exact_and_rk4_match_reference_lsoda(golden):
    g = golden("ode_ref06.npz")
    ex = ode_oracle.exact_solution(ode_oracle.STYLE_REF06, g["y0"], g["rates"], 20.0, 20)
    assert np.abs(ex - g["traj"]).max() < 2e-7            # LSODA @1.49e-8 vs closed form
    # RK4 truncation: err <= 0.01 (h*lam)^4 with lam = largest total outflow rate; rates in
    # this fixture reach lam = 0.95 (all six swept inside the fit bounds), so 8 sub-steps per
    # output interval are needed for <= 1e-6 (4 give 6e-6, 1 gives 4e-3).
    r8 = ode_oracle.rk4(ode_oracle.STYLE_REF06, g["y0"], g["rates"], 20.0, 20, substeps=8)
    assert np.abs(r8 - ex).max() < 4e-7
    assert np.abs(r8 - g["traj"]).max() < 5e-7
    assert np.abs(r8.sum(axis=2) - 1).max() < 1e-15
    k = g["rates"]
    lam = np.maximum(np.maximum(k[0] + k[1], k[2] + k[3]), k[4] + k[5])
    for s in (2, 4, 16):
        e = np.abs(ode_oracle.rk4(ode_oracle.STYLE_REF06, g["y0"], k, 20.0, 20, s) - ex).max(axis=(1, 2))
        assert (e <= 0.0105 * (20.0 / 19 / s * lam) ** 4 + 1e-12).all()


def test_modulated_solve_matches_reference_05(golden):
    """CognitiveStateODE.solve_with_modulation (05:171-196): the LSODA restatement and the node-table RK4 (what the CUDA
    kernel integrates) against the live reference's outputs."""
    g = golden("ode_ref05_modulation.npz")
    base = dict(synth.DEFAULT_RATES)
    for name, fn, y0, t_span, n_points in ode_oracle.modulation_cases():
        t, sol = ode_oracle.solve_with_modulation(base, y0, t_span, fn, n_points)
        assert np.array_equal(t, g[name + "_t"])
        assert np.abs(sol - g[name + "_sol"]).max() < 1e-12, name        # same LSODA, same right-hand side
        S = 8
        tn = np.linspace(t_span[0], t_span[1], 2 * S * (n_points - 1) + 1)
        nodes = np.array([[fn(tt, dict(base))[k] for k in ode_oracle.RATE_ORDER] for tt in tn])
        r = ode_oracle.rk4_modulated(ode_oracle.STYLE_REF06, [y0], nodes, t_span, n_points, S)[0]
        assert np.abs(r - g[name + "_sol"]).max() < 2e-7, name           # LSODA's own error is ~4e-8
        assert np.abs(r.sum(axis=1) - 1).max() < 1e-15


def test_rk45_restatement_matches_reference_solve_ivp(golden):
    g = golden("ode_ref05.npz")
    out, stats = ode_oracle.rk45_scipy(ode_oracle.STYLE_REF06, g["y0"], g["rates"], 20.0, 20, return_stats=True)
    assert np.abs(out - g["traj_rk45"]).max() < 1e-13     # same algorithm, same decisions
    assert all(s["accepted"] >= 3 for s in stats)
    ex = ode_oracle.exact_solution(ode_oracle.STYLE_REF06, g["y0"][:8], g["rates"][:, :8], 50.0, 100)
    assert np.abs(ex - g["traj_odeint_100"]).max() < 2e-7


def test_forecast_style_matches_reference_08(golden):
    g = golden("ode_ref08.npz")
    y0 = ode_oracle.prob_to_state_08(g["p_closed"])
    assert np.abs(y0 - g["y0"]).max() < 1e-15
    n = len(y0)
    rates = np.where((np.arange(n) % 2 == 0)[None, :], g["rates_default"][:, None], g["rates_fit"][:, None])
    ex = ode_oracle.exact_solution(ode_oracle.STYLE_REF08, y0, rates, 20.0, 21)
    assert np.abs(ex - g["traj"]).max() < 2e-7
    r4 = ode_oracle.rk4(ode_oracle.STYLE_REF08, y0, rates, 20.0, 21, substeps=8)
    assert np.abs(r4 - g["traj"]).max() < 3e-7
    # multistep_forecast read-out (08:252-289)
    probs = g["series_probs"]
    m = len(probs) - 20
    y0s = ode_oracle.prob_to_state_08(probs[:m, 1])
    tr = ode_oracle.exact_solution(ode_oracle.STYLE_REF08, y0s, g["rates_default"], 20.0, 21)
    pred = ode_oracle.forecast_readout_08(tr, (5, 10, 20))
    assert np.abs(pred - g["fc_pred"]).max() < 2e-7
    actual = np.stack([probs[h:h + m, 1] for h in (5, 10, 20)], axis=1)
    assert np.array_equal(actual, g["fc_actual"])


def test_pipeline_matches_reference_predict_batch(golden):
    g = golden("pipeline_h128.npz")
    gl = golden("lstm_h128.npz")
    params, x = _lstm_inputs(gl)
    probs, attn = torch_port.forward_probs(torch_port.build_port(params), x, batch_size=3)
    assert np.abs(probs - g["probs"]).max() < 1e-6
    k = ode_oracle.modulate_rates(ode_oracle.rates_to_array(synth.DEFAULT_RATES), 0.5, g["probs"][:, 1], g["probs"][:, 0])
    y0 = ode_oracle.initial_state_06(g["probs"][:, 0], g["probs"][:, 1])
    tr = ode_oracle.exact_solution(ode_oracle.STYLE_REF06, y0, k, 20.0, 20)
    assert np.abs(tr - g["traj"]).max() < 2e-7
    assert np.array_equal(ode_oracle.final_prediction_06(tr), g["preds"])
    assert np.abs(tr[:, -1] - g["three_state"]).max() < 2e-7
    assert np.array_equal(ode_oracle.three_state_class_10(tr[:, -1]), g["cls"])
    # predict_trajectory(X[:1], forecast_steps=10): t = linspace(0,10,10) (06:299-301)
    tr1 = ode_oracle.exact_solution(ode_oracle.STYLE_REF06, y0[:1], k[:, :1], 10.0, 10)
    assert np.abs(tr1[0] - g["single_traj"]).max() < 2e-7


# ---- preprocessing (SURVEY.md §8 f row 4): oracle vs the live reference's 02_preprocessing.py outputs -------------------
def test_preproc_oracle_matches_reference_golden(golden):
    from oracle import preproc_oracle as po
    from lstm_ode_bci_b200 import synth
    g = golden("preproc_ref02.npz")
    seed, C, n = int(g["seed"]), int(g["C"]), int(g["n"])
    raw = synth.make_raw_eeg(seed, 1, C, n)[0]
    filt = po.bandpass_filter(raw, 1.0, 45.0, 500, 4)
    assert np.abs(filt[::7] - g["filtered"]).max() <= 1e-12 * np.abs(g["filtered"]).max()
    X, y, prm = po.preprocess_recording(raw, 1)
    assert X.shape == (int(g["n_seq"]), 256, C) and np.array_equal(y, g["y"])
    assert np.abs(np.asarray(prm["mean"]) - g["mean"]).max() <= 1e-18 and np.abs(np.asarray(prm["std"]) / g["std"] - 1).max() <= 1e-12
    assert np.abs(X.astype(np.float32)[:, :, ::5] - g["X"]).max() <= 1e-6
    raw2 = synth.make_raw_eeg(seed + 1, 1, C, n)[0]
    X2, _, _ = po.preprocess_recording(raw2, 0, prm)
    assert np.abs(X2.astype(np.float32)[::3, :, ::9] - g["X2"]).max() <= 1e-6


def test_preproc_recursion_restatement_matches_scipy():
    """The plain-numpy DF2T recursion (what the CUDA kernel implements) == scipy.signal.filtfilt on a short signal."""
    from scipy.signal import filtfilt
    from oracle import preproc_oracle as po
    rng = np.random.default_rng(3)
    x = rng.standard_normal((3, 400)) + 5.0
    b, a = po.butter_band(1.0, 45.0, 500, 4)
    want = filtfilt(b, a, x, axis=1)
    got = po.filtfilt_restated(b, a, x, use_numpy_recursion=True)
    assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()
    with pytest.raises(ValueError):
        po.filtfilt_restated(b, a, x[:, :27])


# ---- ablation variants (SURVEY.md §8 f row 3): oracle / port vs the live reference's AblationLSTMModel -----------------
ABLATION_TAGS = ["full", "noattn", "unidir", "layers1", "layers2", "minimal256", "unidir256", "noln", "noln_unidir_mean"]


def ablation_case(g, tag):
    from lstm_ode_bci_b200 import synth
    sw, sx, H, L, bidir, att, ln, B, T, C = (int(v) for v in g[tag + ":cfg"])
    params = synth.make_lstm_params(sw, C, H, L, bidirectional=bool(bidir), logit_gain=4.0, use_attention=bool(att), use_layer_norm=bool(ln))
    x = synth.make_windows(sx, B, T, C)
    y = (np.arange(B) % 2).astype(np.int64)
    return params, x, y


@pytest.mark.parametrize("tag", ABLATION_TAGS)
def test_ablation_oracles_match_reference_golden(golden, tag):
    from oracle import lstm_oracle, torch_port
    g = golden("ablation_ref09.npz")
    params, x, y = ablation_case(g, tag)
    logits, attn = lstm_oracle.forward(params, x)
    assert np.abs(logits - g[tag + ":logits"]).max() <= 2e-6
    assert np.abs(attn.sum(axis=1) - 1).max() <= 1e-12
    port = torch_port.build_port(params, dropout=0.0)
    loss, grads, dx, lg = torch_port.loss_and_grads(port, x, y)
    assert np.abs(lg - g[tag + ":logits"]).max() <= 1e-6 and abs(loss - float(g[tag + ":loss"])) <= 1e-6
    for k, gr in grads.items():
        assert abs(np.linalg.norm(gr.astype(np.float64)) - float(g[tag + ":gnorm:" + k])) <= 1e-5 * max(float(g[tag + ":gnorm:" + k]), 1e-3), k
    assert abs(np.linalg.norm(dx.astype(np.float64)) - float(g[tag + ":dx_norm"])) <= 1e-5 * float(g[tag + ":dx_norm"])


def test_fit_objective_matches_reference_seeded_fit(golden):
    """Pins the oracle's restatement of fit_to_data's objective (05_ode_model.py:259-283: MSE + 1e-3 |k|^2 over the clipped,
    renormalised solution) on the reference's own seeded fit result: objective(reference's fitted rates) == reference's loss."""
    from oracle import ode_oracle as oo
    g = golden("ode_ref05_fit.npz")
    obs, tp, k = g["observed"], g["time_points"], g["fitted"]
    sol = oo.exact_solution(oo.STYLE_REF06, obs[:1], k[:, None], float(tp[-1] - tp[0]), len(tp))[0]
    loss = float(np.mean((sol - obs) ** 2) + 0.001 * np.sum(k ** 2))
    assert abs(loss - float(g["loss"])) <= 1e-9, (loss, float(g["loss"]))
    # and the fit really is a (bound-constrained) minimum of it: nudging any rate inside its bounds does not lower the loss
    bounds = [(0.01, 0.5), (0.001, 0.2), (0.02, 0.5), (0.01, 0.3), (0.01, 0.3), (0.02, 0.4)]      # 05:287-294
    for i, (lo, hi) in enumerate(bounds):
        for d in (-1e-3, 1e-3):
            kk = k.copy()
            kk[i] = min(max(kk[i] + d, lo), hi)
            s2 = oo.exact_solution(oo.STYLE_REF06, obs[:1], kk[:, None], float(tp[-1] - tp[0]), len(tp))[0]
            assert float(np.mean((s2 - obs) ** 2) + 0.001 * np.sum(kk ** 2)) >= loss - 1e-9


def explain_case_inputs(g):
    """Model parameters, test windows and labels of tests/golden/explain_ref07.npz, regenerated from its stored seeds
    (tests/golden/make_golden_explain.py: perm_case_params)."""
    params = synth.make_lstm_params(int(g["param_seed"]), 61, 128, 3, logit_gain=float(g["logit_gain"]))
    params["input_proj.0.weight"][:, g["boosted_channels"]] *= np.float32(g["boost"])
    params["classifier.6.bias"] = g["final_bias"].astype(np.float32)
    X = synth.make_windows(int(g["x_seed"]), int(g["n_test"]), 256, 61, structured=True)
    return params, X, g["labels"]


def test_permutation_importance_oracle_matches_reference_07(golden):
    """oracle/explain_oracle.py (07:287-361 restated around the torch port) against the live reference's own run, seeded the same
    way: same subset, same permutations, same predictions -> the same importance per channel, exactly."""
    import torch
    from oracle import explain_oracle
    g = golden("explain_ref07.npz")
    params, X, y = explain_case_inputs(g)
    port = torch_port.build_port(params, dropout=0.0).eval()
    with torch.no_grad():
        assert np.abs(port(torch.from_numpy(X)).numpy() - g["logits"]).max() <= 1e-4      # gain-200 logits of O(10)

    def predict(Xb):
        with torch.no_grad():
            return port(torch.from_numpy(np.ascontiguousarray(Xb))).argmax(dim=1).numpy()

    np.random.seed(int(g["numpy_seed"]))
    imp, base = explain_oracle.permutation_importance(predict, X, y, int(g["n_permutations"]), int(g["n_samples"]))
    assert np.array_equal(imp, g["importance"])
    assert imp.max() > 0.1 and (imp != 0).sum() > 20            # the case discriminates channels
    order = np.argsort(-imp, kind="stable")
    assert set(order[:3]) == set(g["sorted_order"][:3])


def test_sensitivity_oracle_matches_reference_05(golden):
    """05:687-719 restated around the exact steady state (null vector of the generator matrix) against the live reference's
    sensitivity_analysis / get_steady_state (LSODA to t = 1000): 1e-7 on the state, 2e-5 on the finite differences (the
    reference's own integration error, amplified by 1 / (0.4 k))."""
    from oracle import explain_oracle
    g = golden("ode_ref05_sensitivity.npz")

    def steady(p):
        Q = ode_oracle.generator_matrix(ode_oracle.rates_to_array(p))      # dy/dt = Q^T y (05:236-240): pi spans null(Q^T)
        w, v = np.linalg.eig(Q.T)
        pi = np.real(v[:, np.argmin(np.abs(w))])
        return pi / pi.sum()

    for j in (0, 1):
        base = {k: float(g["params_%d" % j][i]) for i, k in enumerate(synth.RATE_ORDER)}
        ss = steady(base)
        assert np.abs(ss - g["steady_%d" % j]).max() <= 1e-7
        assert list(g["names_%d" % j]) == list(base)
        sens = explain_oracle.sensitivity(base, steady)
        assert np.abs(sens - g["sens_%d" % j]).max() <= 2e-5, np.abs(sens - g["sens_%d" % j]).max()
