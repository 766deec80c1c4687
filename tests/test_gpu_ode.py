"""GPU parity of the ODE ensemble kernel (through the C ABI) against the golden vectors of the
live reference and against the CPU oracle.  Tolerance (north_star): <= 1e-6 max-abs on A/P/F with
A+P+F = 1 conserved."""
import numpy as np
import pytest
import torch

from lstm_ode_bci_b200 import ode, ops, synth, integration
from oracle import ode_oracle as oo

pytestmark = pytest.mark.gpu
TOL = 1e-6


def _rates_soa(base64):
    return torch.tensor(base64, dtype=torch.float32)


def test_rk4_coupled_matches_reference_lsoda(golden):
    g = golden("ode_ref06.npz")
    n = g["traj"].shape[0]
    for substeps in (0, 16):
        traj, final, steps = ode.solve_ensemble(n, p_open=g["p_open"], p_closed=g["p_closed"], rates=_rates_soa(g["base"]),
                                                alpha_arr=g["alpha"].astype(np.float32), y0_mode="probs06", coupling=True,
                                                style="ref06", mode="rk4", t_end=20.0, n_points=20, substeps=substeps,
                                                want_steps=True)
        t = traj.cpu().numpy().astype(np.float64)
        assert np.abs(t - g["traj"]).max() <= TOL, substeps
        assert np.abs(t.sum(axis=2) - 1.0).max() <= 2e-7          # fp32 output rounding only
        assert (t >= 0).all()
        assert np.array_equal(final.cpu().numpy(), traj[:, -1].cpu().numpy())
        assert (steps.cpu().numpy() >= 19).all()
    # explicit 8 sub-steps: truncation 3.3e-7 at the stiffest case of this fixture (see oracle test)
    traj8, _, _ = ode.solve_ensemble(n, p_open=g["p_open"], p_closed=g["p_closed"], rates=_rates_soa(g["base"]),
                                     alpha_arr=g["alpha"].astype(np.float32), y0_mode="probs06", coupling=True,
                                     t_end=20.0, n_points=20, substeps=8)
    assert np.abs(traj8.cpu().numpy() - g["traj"]).max() <= TOL


def test_rk4_matches_fp64_oracle_same_method(golden):
    g = golden("ode_ref06.npz")
    n = g["traj"].shape[0]
    want = oo.rk4(oo.STYLE_REF06, g["y0"], g["rates"], 20.0, 20, substeps=8)
    y0 = torch.tensor(g["y0"].T.copy(), dtype=torch.float32)
    traj, _, _ = ode.solve_ensemble(n, rates=torch.tensor(g["rates"], dtype=torch.float32), y0=y0, y0_mode="given",
                                    coupling=False, t_end=20.0, n_points=20, substeps=8)
    assert np.abs(traj.cpu().numpy() - want).max() <= 3e-7        # fp32 arithmetic vs fp64, same scheme


def test_rk45_matches_reference_solve_ivp(golden):
    g = golden("ode_ref05.npz")
    n = g["traj_rk45"].shape[0]
    y0 = torch.tensor(g["y0"].T.copy(), dtype=torch.float32)
    # y0 travels as fp32 over the ABI; use the fp32-rounded y0 for the oracle comparison too
    y0_32 = y0.numpy().T.astype(np.float64)
    want = oo.rk45_scipy(oo.STYLE_REF06, y0_32, g["rates"].astype(np.float32).astype(np.float64), 20.0, 20)
    for f64 in (True, False):
        traj, final, steps = ode.solve_ensemble(n, rates=torch.tensor(g["rates"], dtype=torch.float32), y0=y0,
                                                y0_mode="given", coupling=False, style="ref06", mode="rk45", t_end=20.0,
                                                n_points=20, rtol=1e-3, atol=1e-6, f64=f64, want_steps=True)
        t = traj.cpu().numpy().astype(np.float64)
        assert np.abs(t - want).max() <= (1e-10 if f64 else 1e-7)
        assert np.abs(t - g["traj_rk45"]).max() <= TOL          # vs the reference's own RK45 output
        assert (steps.cpu().numpy() >= 3).all()


def test_forecast_style_matches_reference_08(golden):
    g = golden("ode_ref08.npz")
    n = len(g["p_closed"])
    rates = np.where((np.arange(n) % 2 == 0)[None, :], g["rates_default"][:, None], g["rates_fit"][:, None])
    traj, _, _ = ode.solve_ensemble(n, p_closed=g["p_closed"], rates=torch.tensor(rates, dtype=torch.float32),
                                    y0_mode="pclosed08", coupling=False, style="ref08", mode="rk4", t_end=20.0,
                                    n_points=21, substeps=16)
    t = traj.cpu().numpy()
    assert np.array_equal(t[:, 0].astype(np.float64), g["y0"].astype(np.float32).astype(np.float64))
    assert np.abs(t - g["traj"]).max() <= TOL
    # mirrors of the 08 functions
    tr1 = integration.predict_trajectory(g["y0"][5], synth.DEFAULT_RATES, 10)
    assert tr1.shape == (11, 3) and tr1.dtype == np.float64
    assert np.abs(tr1 - g["traj_n10"][5]).max() <= TOL
    fc = integration.multistep_forecast(g["series_probs"], synth.DEFAULT_RATES, horizons=[5, 10, 20])
    for j, h in enumerate((5, 10, 20)):
        assert np.abs(fc[h]["predictions"] - g["fc_pred"][:, j]).max() <= TOL
        assert np.array_equal(fc[h]["actuals"], g["fc_actual"][:, j])
    roll = integration.rolling_forecast_evaluation(g["series_probs"], synth.DEFAULT_RATES, window_size=10, horizon=10)
    assert np.allclose(roll["mae"].to_numpy(), g["roll_mae"], atol=1e-6)
    assert np.array_equal(roll["accuracy"].to_numpy(), g["roll_accuracy"])


def test_cognitive_state_ode_object_contract(golden):
    g = golden("ode_ref05.npz")
    prm = {k: float(g["rates"][j, 3]) for j, k in enumerate(synth.RATE_ORDER)}
    m = ode.CognitiveStateODE(dict(prm))
    t, sol = m.solve(list(g["y0"][3]), (0, 50), 100)
    assert t.shape == (100,) and sol.shape == (100, 3) and sol.dtype == np.float64
    assert np.abs(sol - g["traj_odeint_100"][3]).max() <= TOL
    t, sol = m.solve(list(g["y0"][3]), (0, 20), 20, method="solve_ivp")
    assert np.abs(sol - g["traj_rk45"][3]).max() <= TOL
    ss = ode.CognitiveStateODE().get_steady_state()          # n_points = 1000: chunked staging path
    Q = ode.CognitiveStateODE().get_transition_matrix()
    pi = np.array([ss["Active"], ss["Passive"], ss["Fatigued"]])
    assert np.abs(pi @ Q).max() < 1e-6 and abs(pi.sum() - 1) < 1e-6


def test_solve_with_modulation_matches_reference_05(golden):
    """05:171-196 through the drop-in object (host samples the callback at the RK4 stage times, one launch integrates)."""
    g = golden("ode_ref05_modulation.npz")
    for name, fn, y0, t_span, n_points in oo.modulation_cases():
        m = ode.CognitiveStateODE()
        t, sol = m.solve_with_modulation(y0, t_span, fn, n_points=n_points)
        assert sol.shape == (n_points, 3) and sol.dtype == np.float64
        assert np.array_equal(t, g[name + "_t"])
        assert np.abs(sol - g[name + "_sol"]).max() <= 1e-6, (name, np.abs(sol - g[name + "_sol"]).max())
        assert np.abs(sol.sum(axis=1) - 1).max() <= 1e-15
    # identity modulation == plain solve
    t, a = ode.CognitiveStateODE().solve_with_modulation([0.2, 0.2, 0.6], (5.0, 25.0), lambda t, p: p, n_points=20)
    _, b = ode.CognitiveStateODE().solve([0.2, 0.2, 0.6], (5.0, 25.0), 20)
    assert np.abs(a - b).max() <= 1e-6


def test_modulated_ensemble_matches_oracle_same_method():
    """bci_ode_solve_modulated against the fp64 numpy restatement of the same node-table RK4: shared and per-trajectory
    schedules, both styles, ragged N."""
    rng = np.random.default_rng(5)
    n, n_points, S, t_span = 333, 20, 4, (0.0, 20.0)
    tn = ode.modulation_nodes(*t_span, n_points, S)
    base = np.array([synth.DEFAULT_RATES[k] for k in synth.RATE_ORDER])
    y0 = rng.dirichlet([2, 2, 2], size=n)
    shared = base[None, :] * (1.0 + 0.5 * np.sin(0.3 * tn)[:, None] * np.array([0, 1, 0, 1, -1, 0])[None, :])
    per = shared[:, :, None] * rng.uniform(0.5, 1.5, size=(1, 6, n))
    for nodes in (shared, per):
        for style, st in (("ref06", oo.STYLE_REF06), ("ref08", oo.STYLE_REF08)):
            traj, fin = ode.solve_modulated_ensemble(y0.T.copy(), nodes, t_span, n_points, S, style=style)
            want = oo.rk4_modulated(st, y0, nodes, t_span, n_points, S)
            got = traj.cpu().numpy()
            assert np.abs(got - want).max() <= 1e-13
            assert np.array_equal(fin.cpu().numpy(), got[:, -1])
    with pytest.raises(Exception):
        ode.solve_modulated_ensemble(y0.T.copy(), shared[:-1], t_span, n_points, S)


@pytest.mark.parametrize("n", [1, 127, 128, 129, 10_000])
def test_sweep_matches_oracle_and_invariants(n):
    sw = synth.make_ode_sweep(42, n)
    traj, final, _ = ode.solve_ensemble(n, p_open=sw["p_open"], p_closed=sw["p_closed"], rates=sw["rates"],
                                        alpha_arr=sw["alpha"], y0_mode="probs06", coupling=True, t_end=20.0, n_points=20,
                                        substeps=8)
    k = oo.modulate_rates(sw["rates"].astype(np.float64), sw["alpha"], sw["p_closed"], sw["p_open"])
    y0 = oo.initial_state_06(sw["p_open"], sw["p_closed"])
    want = oo.rk4(oo.STYLE_REF06, y0, k, 20.0, 20, substeps=8)
    t = traj.cpu().numpy()
    assert np.abs(t - want).max() <= 3e-7
    m = min(n, 128)
    ex = oo.exact_solution(oo.STYLE_REF06, y0[:m], k[:, :m], 20.0, 20)
    assert np.abs(t[:m] - ex).max() <= TOL
    pred, cls = ops.ode_classify(final)
    assert np.array_equal(pred.cpu().numpy(), oo.final_prediction_06(want))
    assert np.array_equal(cls.cpu().numpy(), oo.three_state_class_10(want[:, -1]))


def test_large_ensemble_properties():
    n = 1 << 20
    sw = synth.make_ode_sweep(7, n)
    dev = {k: torch.tensor(v).cuda() for k, v in sw.items()}
    traj, final, _ = ode.solve_ensemble(n, p_open=dev["p_open"], p_closed=dev["p_closed"], rates=dev["rates"],
                                        alpha_arr=dev["alpha"], y0_mode="probs06", coupling=True, substeps=8)
    s = traj.double().sum(dim=2)                                  # fp64 sum: measure the outputs, not the summation
    assert float((s - 1).abs().max()) <= 2e-7                     # A+P+F = 1 to fp32 output rounding
    assert float(traj.min()) >= 0.0 and float(traj.max()) <= 1.0
    assert torch.equal(final, traj[:, -1])
    # alpha = 0  <=>  no coupling at all
    a0, _, _ = ode.solve_ensemble(n, p_open=dev["p_open"], p_closed=dev["p_closed"], rates=dev["rates"], alpha=0.0,
                                  y0_mode="probs06", coupling=True, substeps=8, want_traj=False)[0:3]
    _, f0, _ = ode.solve_ensemble(n, p_open=dev["p_open"], p_closed=dev["p_closed"], rates=dev["rates"], alpha=0.0,
                                  y0_mode="probs06", coupling=True, substeps=8, want_traj=False)
    _, f1, _ = ode.solve_ensemble(n, p_open=dev["p_open"], p_closed=dev["p_closed"], rates=dev["rates"],
                                  y0_mode="probs06", coupling=False, substeps=8, want_traj=False)
    assert torch.equal(f0, f1)
    # more fatigue pressure (larger P(closed)) never lowers the final Fatigued share, all else equal
    pc = torch.linspace(0.61, 1.0, 4096, device="cuda")
    _, ff, _ = ode.solve_ensemble(4096, p_open=1 - pc, p_closed=pc, alpha=1.0, y0_mode="probs06", coupling=True,
                                  substeps=8, want_traj=False)
    assert bool((ff[1:, 2] - ff[:-1, 2] >= -1e-6).all())


def test_edge_cases_and_errors():
    from lstm_ode_bci_b200 import _native
    tr, fs, _ = ode.solve_ensemble(0, p_open=torch.zeros(0), p_closed=torch.zeros(0), y0_mode="probs06", coupling=True)
    assert tr.shape == (0, 20, 3) and fs.shape == (0, 3)
    tr, _, _ = ode.solve_ensemble(3, y0=torch.tensor([[1., 0, 0], [0, 1, 0], [0, 0, 1]]).T.contiguous(), n_points=2, t_end=1.0)
    assert tr.shape == (3, 2, 3)
    assert np.abs(tr.cpu().numpy()[:, 0] - np.eye(3)).max() == 0
    # rate floor 0.001 (06:262): zero base rates are lifted to the floor => state still moves
    tr, _, _ = ode.solve_ensemble(1, p_open=torch.tensor([0.3]), p_closed=torch.tensor([0.7]), base_rates=[0.0] * 6,
                                  y0_mode="probs06", coupling=True, substeps=4)
    want = oo.exact_solution(oo.STYLE_REF06, [[0.2, 0.2, 0.6]], np.full((6, 1), 0.001), 20.0, 20)
    assert np.abs(tr.cpu().numpy() - want).max() <= TOL
    with pytest.raises(_native.BciError):
        ode.solve_ensemble(4, y0_mode="probs06", coupling=True)        # missing probabilities
    with pytest.raises(_native.BciError):
        ode.solve_ensemble(4, y0=torch.zeros(3, 4), n_points=1)


def test_packed_two_trajectory_kernel_is_bit_identical_to_scalar(tmp_path):
    """ode_rk4x2_kernel (two trajectories per thread on FFMA2/FMUL2/FADD2) performs, per lane, exactly the scalar kernel's
    operation sequence: trajectories, final states and both styles agree BIT FOR BIT, including a ragged last block (odd n:
    the last thread's second lane is dead) and the final-state-only launch."""
    import os, subprocess, sys
    code = (
        "import sys, numpy as np, torch\n"
        "from lstm_ode_bci_b200 import ode, synth\n"
        "out = {}\n"
        "for n in (1, 129, 100003):\n"
        "    sw = synth.make_ode_sweep(7, n)\n"
        "    dev = {k: torch.from_numpy(v).cuda() for k, v in sw.items()}\n"
        "    for style, y0m in (('ref06', 'probs06'), ('ref08', 'pclosed08')):\n"
        "        for S in (1, 8):\n"
        "            t, f, _ = ode.solve_ensemble(n, p_open=dev['p_open'], p_closed=dev['p_closed'], rates=dev['rates'], alpha_arr=dev['alpha'],\n"
        "                                         y0_mode=y0m, coupling=True, style=style, mode='rk4', t_end=20.0, n_points=20, substeps=S)\n"
        "            _, f2, _ = ode.solve_ensemble(n, p_open=dev['p_open'], p_closed=dev['p_closed'], rates=dev['rates'], alpha_arr=dev['alpha'],\n"
        "                                          y0_mode=y0m, coupling=True, style=style, mode='rk4', t_end=20.0, n_points=20, substeps=S, want_traj=False)\n"
        "            assert torch.equal(f, f2) and torch.equal(t[:, -1], f)\n"
        "            out['%d_%s_%d' % (n, style, S)] = t.cpu().numpy()\n"
        "np.savez(sys.argv[1], **out)\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for tag, env in (("packed", {}), ("scalar", {"BCI_ODE_RK4": "scalar"})):
        path = str(tmp_path / (tag + ".npz"))
        r = subprocess.run([sys.executable, "-c", code, path], env=dict(os.environ, PYTHONPATH=root, **env), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        res[tag] = dict(np.load(path))
    assert set(res["packed"]) == set(res["scalar"]) and len(res["packed"]) == 12
    for k in res["packed"]:
        assert np.array_equal(res["packed"][k], res["scalar"][k]), k
        assert np.isfinite(res["packed"][k]).all()
