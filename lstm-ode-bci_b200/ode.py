"""Host-side mirror of the reference's three-state ODE model (05_ode_model.py:58-346,
copies at 06:146-180, 10:117-151; functional form 08:132-153).

`CognitiveStateODE` keeps the reference's object contract -- a mutable `.params` dict that
callers reassign per sample (06:296,304,386,404) and `solve(initial_state, t_span, n_points
[, method]) -> (t, solution)` returning float64 numpy -- but integrates on the GPU.  The batched
entry point `solve_ensemble` is what the pipeline mirrors use: one launch for N trajectories.
"""
import numpy as np
import torch

from . import _native as N
from . import ops
from .synth import DEFAULT_RATES, RATE_ORDER


def _dev(device):
    d = torch.device(device if device is not None else "cuda")
    if d.type != "cuda":
        raise N.BciError(-1, "bci_b200 ODE solver runs on CUDA only (no CPU fallback)")
    return torch.device("cuda", d.index if d.index is not None else torch.cuda.current_device())


def _as_dev(a, dev, shape=None):
    if a is None:
        return None
    t = torch.as_tensor(a)
    if not t.is_cuda:
        t = t.to(dev, non_blocking=True)
    t = t.to(torch.float32)
    return t.contiguous()


def solve_ensemble(n, *, p_open=None, p_closed=None, rates=None, base_rates=None, alpha=0.5, alpha_arr=None,
                   y0=None, y0_mode="given", coupling=False, style="ref06", mode="rk4", t_end=20.0,
                   n_points=20, substeps=8, rtol=1e-3, atol=1e-6, want_traj=True, want_steps=False,
                   f64=False, device=None):
    """N independent coupled trajectories in one launch (tensors stay on the device).

    rates      (6,N) per-trajectory base rates (k_ap,k_af,k_pa,k_pf,k_fa,k_fp) or None -> base_rates dict/list
    y0         (3,N) SoA initial states when y0_mode == "given"
    y0_mode    "given" | "probs06" (06:377-382) | "pclosed08" (08:215-234)
    style      "ref06": clamp + normalise + clip/renorm (06:174-180) | "ref08": raw (08:149-153)
    mode       "rk4" (fp32, `substeps` per output interval; 0 = automatic) | "rk45" (scipy-exact, fp64)
    Returns (traj (N,n_points,3) or empty, final_state (N,3), n_steps (N) or empty) CUDA tensors.
    """
    dev = _dev(device)
    if base_rates is None:
        base_rates = DEFAULT_RATES
    if isinstance(base_rates, dict):
        base_rates = [float(base_rates[k]) for k in RATE_ORDER]
    y0m = {"given": N.Y0_GIVEN, "probs06": N.Y0_FROM_PROBS_06, "pclosed08": N.Y0_FROM_PCLOSED_08}[y0_mode]
    return ops.ode_ensemble(_as_dev(p_open, dev), _as_dev(p_closed, dev), _as_dev(rates, dev), _as_dev(alpha_arr, dev),
                            _as_dev(y0, dev), [float(v) for v in base_rates], float(alpha), int(n),
                            {"rk4": N.ODE_RK4, "rk45": N.ODE_RK45}[mode],
                            {"ref06": N.STYLE_REF06, "ref08": N.STYLE_REF08}[style], y0m, bool(coupling),
                            float(t_end), int(n_points), int(substeps), float(rtol), float(atol),
                            bool(want_traj), bool(want_steps), bool(f64), dev.index)


def modulation_nodes(t0, t1, n_points, substeps):
    """Times at which a fixed-step RK4 with `substeps` steps per output interval evaluates its stages:
    node m at t0 + m*h/2, M = 2*substeps*(n_points-1) + 1 (include/bci_b200.h: bci_ode_solve_modulated)."""
    m = 2 * int(substeps) * (int(n_points) - 1) + 1
    return np.linspace(float(t0), float(t1), m)


def solve_modulated_ensemble(y0, rate_nodes, t_span, n_points, substeps, style="ref06", want_traj=True, device=None):
    """N trajectories with time-varying rates in one launch (fp64; 05_ode_model.py:171-196 batched).

    y0          (3,N) initial states
    rate_nodes  (M,6) shared schedule or (M,6,N) per-trajectory schedules sampled at modulation_nodes(...)
    Returns (traj (N,n_points,3) float64 or empty, final_state (N,3) float64) CUDA tensors."""
    import ctypes as C
    dev = _dev(device)
    y0 = torch.as_tensor(y0, dtype=torch.float64).to(dev).contiguous()
    nodes = torch.as_tensor(rate_nodes, dtype=torch.float64).to(dev).contiguous()
    n = int(y0.shape[1])
    m = 2 * int(substeps) * (int(n_points) - 1) + 1
    if y0.shape[0] != 3 or nodes.shape[0] != m or nodes.shape[1] != 6 or (nodes.dim() == 3 and nodes.shape[2] != n) \
            or nodes.dim() not in (2, 3):
        raise N.BciError(-1, "solve_modulated_ensemble: y0 must be (3,N) and rate_nodes (%d,6[,N]); got %s and %s"
                         % (m, tuple(y0.shape), tuple(nodes.shape)))
    a = N.OdeModArgs()
    a.n, a.n_points, a.substeps, a.per_trajectory = n, int(n_points), int(substeps), int(nodes.dim() == 3)
    a.style = {"ref06": N.STYLE_REF06, "ref08": N.STYLE_REF08}[style]
    a.t_span = float(t_span[1]) - float(t_span[0])
    traj = torch.empty((n, n_points, 3) if want_traj else (0,), device=dev, dtype=torch.float64)
    final = torch.empty((n, 3), device=dev, dtype=torch.float64)
    a.rate_nodes, a.y0 = nodes.data_ptr(), y0.data_ptr()
    a.traj = traj.data_ptr() if want_traj else None
    a.final_state = final.data_ptr()
    with torch.cuda.device(dev):
        N.check(N.lib().bci_ode_solve_modulated(C.byref(a), torch.cuda.current_stream().cuda_stream))
    return traj, final


class CognitiveStateODE:
    """Drop-in for the reference class (05_ode_model.py:58; 06:146; 10:117)."""

    def __init__(self, params=None, device=None, substeps=0):
        self.params = dict(DEFAULT_RATES) if params is None else params
        self.state_names = ["Active", "Passive", "Fatigued"]
        self.state_labels = ["A", "P", "F"]
        self.device = device
        self.substeps = substeps  # RK4 sub-steps per output interval; 0 = automatic (<= 2e-7 truncation)

    def ode_system(self, y, t, params=None):
        """Right-hand side (05:101-135), host scalar version kept for API completeness."""
        p = self.params if params is None else params
        A, P, F = max(0, y[0]), max(0, y[1]), max(0, y[2])
        return [-p["k_ap"] * A - p["k_af"] * A + p["k_pa"] * P + p["k_fa"] * F,
                p["k_ap"] * A - p["k_pa"] * P - p["k_pf"] * P + p["k_fp"] * F,
                p["k_af"] * A + p["k_pf"] * P - p["k_fa"] * F - p["k_fp"] * F]

    def solve(self, initial_state, t_span, n_points=100, method="odeint"):
        """(t, solution[n_points,3]) float64.  method 'odeint' -> fixed-step RK4 on the GPU accurate
        to the reference's LSODA within 1e-6; 'solve_ivp' -> scipy-exact RK45 (05:137-169)."""
        t0, t1 = float(t_span[0]), float(t_span[1])
        t = np.linspace(t0, t1, n_points)
        y0 = torch.tensor(np.asarray(initial_state, dtype=np.float64).reshape(3, 1), dtype=torch.float32)
        mode = "rk4" if method == "odeint" else "rk45"
        # constant-rate autonomous system: integrating over [t0,t1] == over [0,t1-t0]
        traj, _, _ = solve_ensemble(1, base_rates=self.params, y0=y0, y0_mode="given", coupling=False, style="ref06",
                                    mode=mode, t_end=t1 - t0, n_points=n_points, substeps=self.substeps,
                                    f64=True, device=self.device)
        return t, traj[0].cpu().numpy()

    def solve_with_modulation(self, initial_state, t_span, modulation_func, n_points=100, substeps=None):
        """05:171-196: rates vary with time through `modulation_func(t, params) -> params`.  The reference lets LSODA call it
        at every right-hand-side evaluation; here it is sampled once at the stage times of a fixed-step RK4 (so the stages see
        the exact rates at t, t+h/2, t+h) and the integration is one GPU launch.  `substeps` per output interval: default
        chosen from the sampled rates so that h*lambda <= 0.05 (truncation ~1e-8 for modulations that are smooth on that scale;
        pass a larger value for faster-varying ones).  Returns (t, solution[n_points,3]) float64 like the reference."""
        t0, t1 = float(t_span[0]), float(t_span[1])
        t = np.linspace(t0, t1, n_points)

        def sample(S):
            tn = modulation_nodes(t0, t1, n_points, S)
            return np.array([[float(modulation_func(float(tt), self.params.copy())[k]) for k in RATE_ORDER] for tt in tn])

        if substeps is None:
            coarse = sample(1)
            lam = max(float((coarse[:, 0] + coarse[:, 1]).max()), float((coarse[:, 2] + coarse[:, 3]).max()),
                      float((coarse[:, 4] + coarse[:, 5]).max()), 1e-12)
            substeps = int(min(256, max(2, np.ceil((t1 - t0) / (n_points - 1) * lam / 0.05))))
        nodes = sample(substeps)
        y0 = np.asarray(initial_state, dtype=np.float64).reshape(3, 1)
        traj, _ = solve_modulated_ensemble(y0, nodes, (t0, t1), n_points, substeps, style="ref06", device=self.device)
        return t, traj[0].cpu().numpy()

    FIT_BOUNDS = [(0.01, 0.5), (0.001, 0.2), (0.02, 0.5), (0.01, 0.3), (0.01, 0.3), (0.02, 0.4)]   # 05:287-294

    def population_loss(self, param_matrix, observed_proportions, time_points, substeps=0):
        """Objective of fit_to_data (05:259-283) for a whole population at once: param_matrix (6,S) -> (S,) float64.
        One ODE-ensemble launch integrates all S candidates; MSE + 1e-3 |k|^2 is reduced on the device."""
        obs = torch.as_tensor(np.asarray(observed_proportions, dtype=np.float64))
        S = param_matrix.shape[1]
        dev = _dev(self.device)
        y0 = obs[0].to(torch.float32).reshape(3, 1).repeat(1, S)
        rates = torch.as_tensor(np.ascontiguousarray(param_matrix), dtype=torch.float32)
        traj, _, _ = solve_ensemble(S, rates=rates, y0=y0, y0_mode="given", coupling=False, style="ref06", mode="rk4",
                                    t_end=float(time_points[-1] - time_points[0]), n_points=len(time_points),
                                    substeps=substeps, f64=True, device=dev)
        mse = ((traj - obs.to(dev)[None]) ** 2).mean(dim=(1, 2))
        reg = 0.001 * (rates.to(dev).double() ** 2).sum(dim=0)
        return (mse + reg).cpu().numpy()

    def fit_to_data(self, observed_proportions, time_points, method="differential_evolution"):
        """05:244-322 (SURVEY.md §8 f rank 2).  Same bounds, seed, maxiter, tol, polish and return value; the differential
        evolution evaluates each generation's population in one GPU launch (scipy `vectorized=True`, deferred updating)
        instead of one odeint call per candidate, so the optimiser's path differs from the reference's immediate-updating
        run while converging to the same minimum of the same objective."""
        from scipy.optimize import differential_evolution, minimize
        observed = np.asarray(observed_proportions, dtype=np.float64)
        tp = np.asarray(time_points, dtype=np.float64)

        def pop_loss(x):
            x = np.asarray(x, dtype=np.float64)
            return self.population_loss(x.reshape(6, -1), observed, tp) if x.ndim == 2 else \
                float(self.population_loss(x.reshape(6, 1), observed, tp)[0])

        if method == "differential_evolution":
            result = differential_evolution(pop_loss, self.FIT_BOUNDS, seed=42, maxiter=1000, tol=1e-7, polish=True,
                                            vectorized=True, updating="deferred")
        else:
            result = minimize(pop_loss, [0.1, 0.02, 0.15, 0.08, 0.05, 0.1], bounds=self.FIT_BOUNDS, method="L-BFGS-B",
                              options={"maxiter": 1000})
        fitted = {k: float(result.x[i]) for i, k in enumerate(RATE_ORDER)}
        self.params = fitted
        return fitted, float(result.fun)

    def get_transition_matrix(self):
        p = self.params
        return np.array([[-(p["k_ap"] + p["k_af"]), p["k_ap"], p["k_af"]],
                         [p["k_pa"], -(p["k_pa"] + p["k_pf"]), p["k_pf"]],
                         [p["k_fa"], p["k_fp"], -(p["k_fa"] + p["k_fp"])]])

    def get_steady_state(self):
        """05:198-221: solve to t=1000 with 1000 points and take the last state."""
        _, sol = self.solve([0.33, 0.33, 0.34], (0, 1000), 1000)
        return {"Active": sol[-1][0], "Passive": sol[-1][1], "Fatigued": sol[-1][2]}




def steady_state_ensemble(param_matrix, device=None, substeps=2):
    """get_steady_state (05:198-221: solve from [.33,.33,.34] to t = 1000 over 1000 points, keep the last state) for S parameter sets
    in ONE launch: param_matrix (6,S) in RATE_ORDER -> (S,3) float64 numpy.

    Integrated in fp64 by the node-table RK4 kernel (`bci_ode_solve_modulated` with a constant schedule per trajectory) rather than
    by the fp32 ensemble kernel: callers difference these states (sensitivity_analysis divides a +-20 % difference by 0.4 k, i.e.
    amplifies errors up to 125x), and the fixed point of an RK4 map of a linear system is the exact null vector of Q, so the
    converged fp64 state carries no truncation error at all.  Only the final state is written."""
    pm = np.ascontiguousarray(np.asarray(param_matrix, dtype=np.float64).reshape(6, -1))
    S = pm.shape[1]
    n_points = 1000
    m = 2 * int(substeps) * (n_points - 1) + 1
    nodes = np.broadcast_to(pm[None], (m, 6, S))
    y0 = np.repeat(np.array([[0.33], [0.33], [0.34]], dtype=np.float64), S, axis=1)
    _, final = solve_modulated_ensemble(y0, np.ascontiguousarray(nodes), (0.0, 1000.0), n_points, substeps, style="ref06",
                                        want_traj=False, device=device)
    return final.cpu().numpy()


def sensitivity_analysis(ode_model, output_path=None):
    """05_ode_model.py:687-750 (SURVEY.md §8 f rank 2): central-difference sensitivity of the steady state to each rate,
    +-20 % perturbations -> [{'parameter', 'sens_Active', 'sens_Passive', 'sens_Fatigued'}, ...] in the order of
    `ode_model.params`.  The reference builds 12 models and integrates each to t = 1000 with odeint, one after the other; here the 12
    perturbed parameter sets are one ODE-ensemble launch.  `output_path` is accepted for signature compatibility: the heat-map
    figure (05:721-745) is plotting, outside this library."""
    base_params = dict(ode_model.params)
    param_names = list(base_params.keys())
    perturbation = 0.2
    columns = []
    for param in param_names:
        for factor in (1 - perturbation, 1 + perturbation):
            test_params = dict(base_params)
            test_params[param] = base_params[param] * factor
            columns.append([float(test_params[k]) for k in RATE_ORDER])
    steady = steady_state_ensemble(np.array(columns, dtype=np.float64).T, device=getattr(ode_model, "device", None))
    results = []
    for j, param in enumerate(param_names):
        delta = (steady[2 * j + 1] - steady[2 * j]) / (2 * perturbation * base_params[param])
        results.append({"parameter": param, "sens_Active": delta[0], "sens_Passive": delta[1], "sens_Fatigued": delta[2]})
    return results
