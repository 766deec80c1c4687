"""Generate tests/golden/*.npz from the LIVE reference (build container only).

Run:  python tests/golden/make_golden.py
Needs /root/reference (read-only mount).  The reference has no tests or golden vectors of
its own, so these files -- outputs of the reference's own classes on seeded
inputs -- are what pins the oracle and, through it, the CUDA path.  Inputs are regenerated
from seeds by lstm_ode_bci_b200.synth (numpy PCG64), so only outputs are stored.
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from lstm_ode_bci_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
warnings.filterwarnings("ignore")


def ref_model(mod, params, C, H, L):
    torch.manual_seed(0)
    m = mod.EnhancedLSTMModel(input_size=C, hidden_size=H, num_layers=L, num_classes=2,
                              dropout=0.4, bidirectional=True)
    m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in params.items()}, strict=True)
    m.to("cpu").eval()
    return m


def lstm_case(tag, seed_w, seed_x, C, H, L, B, T, gain, store_inputs=False):
    ref04 = ref_loader.load("ref04")
    params = synth.make_lstm_params(seed_w, C, H, L, logit_gain=gain)
    x = synth.make_windows(seed_x, B, T, C, structured=(seed_x % 2 == 1))
    m = ref_model(ref04, params, C, H, L)
    with torch.no_grad():
        logits, attn = m(torch.from_numpy(x), return_attention=True)
        probs = torch.softmax(logits, dim=1)
    d = dict(seed_w=seed_w, seed_x=seed_x, C=C, H=H, L=L, B=B, T=T, gain=gain,
             structured=int(seed_x % 2 == 1),
             logits=logits.numpy(), probs=probs.numpy(), attention=attn.numpy())
    if store_inputs:
        d["x"] = x
        for k, v in params.items():
            d["param:" + k] = v
    np.savez_compressed(os.path.join(OUT, f"lstm_{tag}.npz"), **d)
    print(tag, "logits[0]", logits[0].numpy(), "probs range", probs.min().item(), probs.max().item())
    return m, params, x


def lstm_grad_case(tag, seed_w, seed_x, C, H, L, B, T):
    """Training-step oracle pin (04_lstm_model.py:486-494, plain CE as in 09:297-303, dropout
    disabled by constructing with dropout=0 so gradients are deterministic)."""
    ref04 = ref_loader.load("ref04")
    params = synth.make_lstm_params(seed_w, C, H, L, logit_gain=4.0)
    x = synth.make_windows(seed_x, B, T, C)
    y = (np.arange(B) % 2).astype(np.int64)
    torch.manual_seed(0)
    m = ref04.EnhancedLSTMModel(input_size=C, hidden_size=H, num_layers=L, num_classes=2,
                                dropout=0.0, bidirectional=True)
    m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in params.items()}, strict=True)
    m.to("cpu").train()
    xb = torch.from_numpy(x).requires_grad_(True)
    w = torch.tensor([0.7, 1.3])
    loss = torch.nn.functional.cross_entropy(m(xb), torch.from_numpy(y), weight=w)
    loss.backward()
    d = dict(seed_w=seed_w, seed_x=seed_x, C=C, H=H, L=L, B=B, T=T, gain=4.0, loss=float(loss),
             class_weight=w.numpy(), y=y, dx_norm=float(xb.grad.norm()),
             dx_head=xb.grad[0, :4, :8].numpy().copy())
    for k, p in m.named_parameters():
        g = p.grad.detach().numpy()
        d["gnorm:" + k] = np.float64(np.linalg.norm(g.astype(np.float64)))
        d["ghead:" + k] = g.reshape(-1)[:16].copy()
        if g.size <= 4096:
            d["gfull:" + k] = g.copy()
    np.savez_compressed(os.path.join(OUT, f"lstm_grad_{tag}.npz"), **d)
    print("grad", tag, "loss", float(loss))


def ode_cases():
    ref05 = ref_loader.load("ref05")
    ref06 = ref_loader.load("ref06")
    ref08 = ref_loader.load("ref08")
    rng = np.random.default_rng(2024)
    n = 96
    sweep = synth.make_ode_sweep(11, n)
    # vary all six base rates inside the fit bounds (05_ode_model.py:287-294) for half the cases
    base = sweep["rates"].astype(np.float64)
    lo = np.array([0.01, 0.001, 0.01, 0.01, 0.01, 0.01])
    hi = np.array([0.5, 0.2, 0.5, 0.3, 0.3, 0.3])
    base[:, n // 2:] = rng.uniform(lo[:, None], hi[:, None], size=(6, n - n // 2))
    base = base.astype(np.float32).astype(np.float64)      # exactly representable in fp32
    p_closed = sweep["p_closed"].copy()
    p_closed[:6] = np.float32([0.0, 1.0, 0.6, 0.4, 0.5, 0.6000001])  # threshold edges
    p_open = (np.float32(1.0) - p_closed).astype(np.float32)
    alpha = sweep["alpha"].astype(np.float64)

    traj06 = np.empty((n, 20, 3))
    rates06 = np.empty((6, n))
    y0_06 = np.empty((n, 3))
    for i in range(n):
        prm = {k: float(base[j, i]) for j, k in enumerate(synth.RATE_ORDER)}
        ode = ref06.CognitiveStateODE(dict(prm))
        integ = ref06.LSTMODEIntegration(None, ode, coupling_strength=float(alpha[i]))
        po, pc = p_open[i], p_closed[i]           # float32 scalars, as predict_batch sees them
        mod = integ.modulate_ode_rates(pc, po)
        rates06[:, i] = [mod[k] for k in synth.RATE_ORDER]
        if pc > 0.6:
            y0 = [0.2, 0.2, 0.6]
        elif po > 0.6:
            y0 = [0.6, 0.2, 0.2]
        else:
            y0 = [0.33, 0.34, 0.33]
        y0_06[i] = y0
        ode.params = mod
        _, traj06[i] = ode.solve(y0, (0, 20), 20)
    np.savez_compressed(os.path.join(OUT, "ode_ref06.npz"), base=base, alpha=alpha, p_open=p_open,
                        p_closed=p_closed, rates=rates06, y0=y0_06, traj=traj06)
    print("ode06 traj[0,-1]", traj06[0, -1])

    # solve_ivp / RK45 branch of 05 (05_ode_model.py:157-163), same couplings (rates06), plus
    # unnormalised initial states to exercise the y0/sum(y0) step
    traj45 = np.empty((n, 20, 3))
    y0_45 = rng.uniform(0.05, 1.0, size=(n, 3))
    y0_45[: n // 2] = y0_06[: n // 2]
    for i in range(n):
        ode = ref05.CognitiveStateODE({k: float(rates06[j, i]) for j, k in enumerate(synth.RATE_ORDER)})
        _, traj45[i] = ode.solve(list(y0_45[i]), (0, 20), 20, method="solve_ivp")
    # and the odeint branch with n_points=100 default horizon (05:137) for a few
    traj05 = np.empty((8, 100, 3))
    for i in range(8):
        ode = ref05.CognitiveStateODE({k: float(rates06[j, i]) for j, k in enumerate(synth.RATE_ORDER)})
        _, traj05[i] = ode.solve(list(y0_45[i]), (0, 50), 100)
    np.savez_compressed(os.path.join(OUT, "ode_ref05.npz"), rates=rates06, y0=y0_45, traj_rk45=traj45,
                        traj_odeint_100=traj05)
    print("ode05 rk45 traj[0,-1]", traj45[0, -1])

    # 08: prob_to_ode_state + predict_trajectory (raw, unmodulated) + readout
    pcs = np.concatenate([np.float32([0.0, 0.5, 0.5000001, 1.0]), rng.uniform(0, 1, 60).astype(np.float32)])
    prm = dict(synth.DEFAULT_RATES)
    prm_fit = {"k_ap": 0.020, "k_af": 0.095, "k_pa": 0.15, "k_pf": 0.626, "k_fa": 0.139, "k_fp": 0.1}  # README.md:228-233
    y0_08 = np.stack([ref08.prob_to_ode_state(p) for p in pcs])
    traj08 = np.stack([ref08.predict_trajectory(y0_08[i], prm if i % 2 == 0 else prm_fit, 20) for i in range(len(pcs))])
    traj08_10 = np.stack([ref08.predict_trajectory(y0_08[i], prm, 10) for i in range(8)])
    # multistep_forecast over a synthetic probability series (08_forecasting.py:252-289)
    series = rng.uniform(0, 1, 64).astype(np.float32)
    probs = np.stack([1 - series, series], axis=1).astype(np.float32)
    import io, contextlib
    ref08.tqdm = lambda it, **k: it
    with contextlib.redirect_stdout(io.StringIO()):
        fc = ref08.multistep_forecast(probs, prm, horizons=[5, 10, 20])
        roll = ref08.rolling_forecast_evaluation(probs, prm, window_size=10, horizon=10)
    np.savez_compressed(os.path.join(OUT, "ode_ref08.npz"), p_closed=pcs, y0=y0_08, traj=traj08, traj_n10=traj08_10,
                        rates_default=np.array([prm[k] for k in synth.RATE_ORDER]),
                        rates_fit=np.array([prm_fit[k] for k in synth.RATE_ORDER]),
                        series_probs=probs,
                        fc_pred=np.stack([fc[h]["predictions"] for h in (5, 10, 20)], axis=1),
                        fc_actual=np.stack([fc[h]["actuals"] for h in (5, 10, 20)], axis=1),
                        roll_accuracy=roll["accuracy"].to_numpy(), roll_mae=roll["mae"].to_numpy())
    print("ode08 traj[1,20]", traj08[1, 20])


def pipeline_cases(m, params, x):
    """predict_batch (06:308-406) and get_three_state_probabilities (10:204-290) end to end."""
    ref06 = ref_loader.load("ref06")
    ref10 = ref_loader.load("ref10")
    import io, contextlib
    m06 = ref_model(ref06, params, 61, 128, 3)
    ode = ref06.CognitiveStateODE()
    integ = ref06.LSTMODEIntegration(m06, ode, coupling_strength=0.5)
    trajs, probs, preds = integ.predict_batch(x, forecast_steps=20, batch_size=3, show_progress=False)
    tr1, pr1, at1 = integ.predict_trajectory(x[:1], forecast_steps=10)
    m10 = ref_model(ref10, params, 61, 128, 3)
    ref10.tqdm = lambda it, **k: it
    with contextlib.redirect_stdout(io.StringIO()):
        lp, three, cls = ref10.get_three_state_probabilities(m10, ref10.CognitiveStateODE(), x, batch_size=4)
    np.savez_compressed(os.path.join(OUT, "pipeline_h128.npz"), traj=trajs, probs=probs, preds=preds,
                        single_traj=tr1, single_probs=pr1, single_attn=at1,
                        lstm_probs10=lp, three_state=three, cls=cls)
    print("pipeline probs", probs[:, 1], "preds", preds, "cls", cls)


if __name__ == "__main__":
    torch.set_num_threads(8)
    lstm_case("tiny", 1, 2, 5, 8, 2, 3, 6, 1.0, store_inputs=True)
    m, params, x = lstm_case("h128", 42, 7, 61, 128, 3, 8, 256, 12.0)
    lstm_case("h128_t64", 43, 8, 61, 128, 3, 5, 64, 1.0)
    lstm_case("h256", 44, 9, 61, 256, 3, 3, 256, 12.0)
    lstm_grad_case("h128", 45, 10, 61, 128, 3, 6, 64)
    lstm_grad_case("tiny", 3, 4, 5, 8, 2, 4, 6)
    ode_cases()
    pipeline_cases(m, params, x)
