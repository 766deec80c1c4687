"""CPU oracle for the three-state A/P/F ODE ensemble with probabilistic rate coupling
(TEST INFRASTRUCTURE ONLY -- never imported by the product path).

Restates, in numpy/float64:
  * rhs                    05_ode_model.py:101-135 == 06_lstm_ode_integration.py:158-172
                           (clamped); 08_forecasting.py:132-146 (unclamped)
  * modulate_rates         06_lstm_ode_integration.py:236-264 (4 modulated rates + 0.001 floor)
  * initial_state_06       06_lstm_ode_integration.py:377-382 == 10_three_state_probabilities.py:250-255
  * prob_to_state_08       08_forecasting.py:215-234
  * post_process_06        06_lstm_ode_integration.py:178-179 (clip[0,1] then row-renormalise)
  * three_state_class_10   10_three_state_probabilities.py:282-288
  * forecast_readout_08    08_forecasting.py:273-282 (F + 0.5 P clipped to [0,1])

The integrators themselves live in third-party scipy (unpinned by the reference:
requirements.txt:10 `scipy>=1.11.0`; 1.18.1 in this image):
  * reference default path = scipy.integrate.odeint (ODEPACK LSODA, rtol=atol=1.49012e-8),
    which is within 3.9e-8 of the exact solution of this linear system (SURVEY.md §8 c).  Its
    oracle here is the closed form y(t) = expm(Q^T t) y0 (`exact_solution`), Q from
    05_ode_model.py:236-240 -- valid because the max(0,.) clamp never activates inside the
    simplex and rates are constant per trajectory.
  * `rk4` is the fixed-step integrator the CUDA kernel implements (our choice of method;
    the reference has no RK4), in float64 so that only truncation error separates it from
    `exact_solution`.
  * `rk45_scipy` restates scipy's RK45 (scipy/integrate/_ivp/rk.py + common.py +
    ivp.py t_eval handling: Dormand-Prince 5(4), select_initial_step, RMS error norm with
    scale = atol + max(|y|,|y_new|) rtol, SAFETY 0.9, MIN/MAX_FACTOR 0.2/10, exponent -1/5,
    no growth after a rejection, quartic dense output at t_eval) -- the non-default branch
    05_ode_model.py:157-163 with scipy's default rtol=1e-3, atol=1e-6.

Parity pin: the reference ships no tests/golden vectors; this oracle is pinned
against outputs of the reference itself run in the build container
(tests/golden/ode_*.npz from tests/golden/make_golden.py) by tests/test_oracle_golden.py.
"""
import numpy as np

RATE_ORDER = ("k_ap", "k_af", "k_pa", "k_pf", "k_fa", "k_fp")
RATE_FLOOR = 0.001

STYLE_REF06 = 0  # clamp in rhs, y0 normalised, t = linspace(t0, t_end, n), clip + renorm
STYLE_REF08 = 1  # raw rhs, y0 as given, t = linspace(0, n*dt, n+1), no post-processing


def rates_to_array(params):
    return np.array([params[k] for k in RATE_ORDER], dtype=np.float64)


def modulate_rates(base, alpha, p_closed, p_open):
    """base (6,) or (6,N); alpha, p_* scalars or (N,).  Returns (6,N) float64.

    dtype note: predict_batch feeds float32 numpy scalars taken from the LSTM's softmax
    (06:373-374) into `params[k] * (1 + alpha * p)`, where params/alpha are Python floats.
    Under NumPy >= 2 promotion rules (Python scalars are weak) that arithmetic happens in
    float32; the result is then used as a float64 rate.  We mirror it: float32 inputs are
    combined in float32 (one rounding per operation), anything else in float64."""
    p_closed = np.atleast_1d(np.asarray(p_closed))
    p_open = np.atleast_1d(np.asarray(p_open))
    wt = np.float32 if (p_closed.dtype == np.float32 and p_open.dtype == np.float32) else np.float64
    p_closed = p_closed.astype(wt)
    p_open = p_open.astype(wt)
    n = p_closed.shape[0]
    base = np.asarray(base, dtype=np.float64)
    k = np.broadcast_to(base.reshape(6, -1), (6, n)).astype(np.float64).copy()
    alpha = np.broadcast_to(np.asarray(alpha, dtype=wt), (n,))
    one = wt(1.0)
    fat = one + alpha * p_closed
    rec = one + alpha * p_open
    k[1] = k[1].astype(wt) * fat  # k_af
    k[3] = k[3].astype(wt) * fat  # k_pf
    k[4] = k[4].astype(wt) * rec  # k_fa
    k[2] = k[2].astype(wt) * rec  # k_pa
    return np.maximum(RATE_FLOOR, k)


def initial_state_06(p_open, p_closed):
    p_open = np.atleast_1d(np.asarray(p_open))
    p_closed = np.atleast_1d(np.asarray(p_closed))
    y0 = np.empty((p_open.shape[0], 3), dtype=np.float64)
    y0[:] = (0.33, 0.34, 0.33)
    y0[p_open > 0.6] = (0.6, 0.2, 0.2)
    y0[p_closed > 0.6] = (0.2, 0.2, 0.6)  # checked first in the reference, so it wins
    return y0


def prob_to_state_08(p_closed):
    """08:215-234.  Like modulate_rates, float32 probabilities (what multistep_forecast passes,
    08:266-267) are combined in float32 under NumPy >= 2 promotion; the float32 result is then
    integrated in float64.  Returns float64 values (of float32 precision in that case)."""
    p = np.atleast_1d(np.asarray(p_closed))
    wt = np.float32 if p.dtype == np.float32 else np.float64
    p = p.astype(wt)
    a = wt(1.0) - p
    hi = p > 0.5
    f = np.where(hi, p * wt(0.6), p * wt(0.3)).astype(wt)
    pp = np.where(hi, p * wt(0.4), p * wt(0.3)).astype(wt)
    tot = (a + pp) + f
    return np.stack([a / tot, pp / tot, f / tot], axis=1).astype(np.float64)


def rhs(y, k, clamp):
    """y (N,3), k (6,N) -> dy (N,3).  Operation order follows the reference expressions."""
    A, P, F = y[:, 0], y[:, 1], y[:, 2]
    if clamp:
        A, P, F = np.maximum(0.0, A), np.maximum(0.0, P), np.maximum(0.0, F)
    k_ap, k_af, k_pa, k_pf, k_fa, k_fp = k
    dA = -k_ap * A - k_af * A + k_pa * P + k_fa * F
    dP = k_ap * A - k_pa * P - k_pf * P + k_fp * F
    dF = k_af * A + k_pf * P - k_fa * F - k_fp * F
    return np.stack([dA, dP, dF], axis=1)


def time_grid(style, t_end, n_points):
    """06:175 linspace(t0, t1, n_points); 08:151 linspace(0, n_steps*dt, n_steps+1) -- the
    caller passes t_end = n_steps*dt and n_points = n_steps+1 for the latter."""
    return np.linspace(0.0, t_end, n_points)


def post_process_06(sol):
    sol = np.clip(sol, 0.0, 1.0)
    return sol / sol.sum(axis=-1, keepdims=True)


def _prepare(style, y0, k):
    y0 = np.asarray(y0, dtype=np.float64).reshape(-1, 3)
    if style == STYLE_REF06:
        y0 = y0 / y0.sum(axis=1, keepdims=True)           # 06:176
    k = np.asarray(k, dtype=np.float64).reshape(6, -1)
    if k.shape[1] == 1 and y0.shape[0] > 1:
        k = np.broadcast_to(k, (6, y0.shape[0]))
    return y0, k


def rk4(style, y0, k, t_end, n_points, substeps):
    """Classical RK4 with `substeps` equal steps per output interval. Returns (N,n_points,3)."""
    y, k = _prepare(style, y0, k)
    clamp = style == STYLE_REF06
    n = y.shape[0]
    out = np.empty((n, n_points, 3), dtype=np.float64)
    out[:, 0] = y
    h = (t_end / (n_points - 1)) / substeps
    for i in range(1, n_points):
        for _ in range(substeps):
            k1 = rhs(y, k, clamp)
            k2 = rhs(y + 0.5 * h * k1, k, clamp)
            k3 = rhs(y + 0.5 * h * k2, k, clamp)
            k4 = rhs(y + h * k3, k, clamp)
            y = y + (h / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
        out[:, i] = y
    return post_process_06(out) if style == STYLE_REF06 else out


def generator_matrix(k):
    """Q of 05_ode_model.py:236-240 for one trajectory; dy/dt = Q^T y."""
    k_ap, k_af, k_pa, k_pf, k_fa, k_fp = k
    return np.array([[-(k_ap + k_af), k_ap, k_af],
                     [k_pa, -(k_pa + k_pf), k_pf],
                     [k_fa, k_fp, -(k_fa + k_fp)]], dtype=np.float64)


def exact_solution(style, y0, k, t_end, n_points):
    """Closed form expm(Q^T t) y0 at the output grid (per trajectory; small N only)."""
    from scipy.linalg import expm
    y0, k = _prepare(style, y0, k)
    t = time_grid(style, t_end, n_points)
    out = np.empty((y0.shape[0], n_points, 3), dtype=np.float64)
    for j in range(y0.shape[0]):
        Qt = generator_matrix(k[:, j]).T
        for i in range(n_points):
            out[j, i] = expm(Qt * t[i]) @ y0[j]
    return post_process_06(out) if style == STYLE_REF06 else out


# ---- scipy-exact Dormand-Prince (see module docstring for the files restated) --------------
_C = np.array([0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1])
_A = [
    [],
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
]
_B = np.array([35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84])
_E = np.array([-71 / 57600, 0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40])
_P = np.array([
    [1, -8048581381 / 2820520608, 8663915743 / 2820520608, -12715105075 / 11282082432],
    [0, 0, 0, 0],
    [0, 131558114200 / 32700410799, -68118460800 / 10900136933, 87487479700 / 32700410799],
    [0, -1754552775 / 470086768, 14199869525 / 1410260304, -10690763975 / 1880347072],
    [0, 127303824393 / 49829197408, -318862633887 / 49829197408, 701980252875 / 199316789632],
    [0, -282668133 / 205662961, 2019193451 / 616988883, -1453857185 / 822651844],
    [0, 40617522 / 29380423, -110615467 / 29380423, 69997945 / 29380423]])


def _rms(x):
    return float(np.sqrt(np.sum(x * x)) / np.sqrt(x.size))


def rk45_scipy_one(f, y0, t_end, t_eval, rtol=1e-3, atol=1e-6):
    """One trajectory, scipy RK45 semantics, t0 = 0, forward.  Returns (len(t_eval),3), stats."""
    t = 0.0
    y = np.array(y0, dtype=np.float64)
    fcur = f(y)
    nfev = 1
    # select_initial_step (common.py)
    scale = atol + np.abs(y) * rtol
    d0 = _rms(y / scale)
    d1 = _rms(fcur / scale)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    h0 = min(h0, t_end)
    f1 = f(y + h0 * fcur)
    nfev += 1
    d2 = _rms((f1 - fcur) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(1e-6, h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1 / 5)
    h_abs = min(100 * h0, h1, t_end)
    out = np.empty((len(t_eval), 3), dtype=np.float64)
    ti = 0
    K = np.empty((7, 3), dtype=np.float64)
    n_acc = n_rej = 0
    while t < t_end:
        min_step = 10 * abs(np.nextafter(t, np.inf) - t)
        h_abs = max(h_abs, min_step)
        rejected = False
        while True:
            if h_abs < min_step:
                raise RuntimeError("step size too small")
            t_new = t + h_abs
            if t_new - t_end > 0:
                t_new = t_end
            h = t_new - t
            h_abs = abs(h)
            K[0] = fcur
            for s in range(1, 6):
                dy = np.dot(K[:s].T, np.array(_A[s])) * h
                K[s] = f(y + dy)
            y_new = y + h * np.dot(K[:6].T, _B)
            f_new = f(y_new)
            nfev += 6
            K[6] = f_new
            sc = atol + np.maximum(np.abs(y), np.abs(y_new)) * rtol
            err = _rms(np.dot(K.T, _E) * h / sc)
            if err < 1:
                factor = 10.0 if err == 0 else min(10.0, 0.9 * err ** -0.2)
                if rejected:
                    factor = min(1.0, factor)
                h_abs *= factor
                n_acc += 1
                break
            h_abs *= max(0.2, 0.9 * err ** -0.2)
            rejected = True
            n_rej += 1
        # dense output at every t_eval <= t_new (ivp.py: searchsorted side='right')
        Q = K.T.dot(_P)
        while ti < len(t_eval) and t_eval[ti] <= t_new:
            x = (t_eval[ti] - t) / h
            p = np.cumprod(np.array([x, x, x, x]))
            out[ti] = h * np.dot(Q, p) + y
            ti += 1
        t, y, fcur = t_new, y_new, f_new
    return out, {"accepted": n_acc, "rejected": n_rej, "nfev": nfev}


def rk45_scipy(style, y0, k, t_end, n_points, rtol=1e-3, atol=1e-6, return_stats=False):
    y0, k = _prepare(style, y0, k)
    clamp = style == STYLE_REF06
    t_eval = time_grid(style, t_end, n_points)
    out = np.empty((y0.shape[0], n_points, 3), dtype=np.float64)
    stats = []
    for j in range(y0.shape[0]):
        kj = k[:, j:j + 1]
        f = lambda y, kj=kj: rhs(y[None, :], kj, clamp)[0]
        out[j], st = rk45_scipy_one(f, y0[j], float(t_end), t_eval, rtol, atol)
        stats.append(st)
    res = post_process_06(out) if style == STYLE_REF06 else out
    return (res, stats) if return_stats else res


def three_state_class_10(final_state):
    """10_three_state_probabilities.py:282-288: 2 if F>.5, 0 if A>.5, else 1."""
    fs = np.asarray(final_state)
    cls = np.ones(fs.shape[0], dtype=np.int64)
    cls[fs[:, 0] > 0.5] = 0
    cls[fs[:, 2] > 0.5] = 2   # checked first in the reference, so it wins
    return cls


def final_prediction_06(traj):
    """06_lstm_ode_integration.py:396-401: 1 if Fatigued > 0.5 at the last point else 0."""
    return (np.asarray(traj)[:, -1, 2] > 0.5).astype(np.int64)


def forecast_readout_08(traj, horizons):
    """08_forecasting.py:273-279: clip(F_h + 0.5 P_h, 0, 1) for each horizon -> (N,len(h))."""
    traj = np.asarray(traj)
    return np.stack([np.clip(traj[:, h, 2] + traj[:, h, 1] * 0.5, 0.0, 1.0) for h in horizons], axis=1)


# ---- the reference's own call pattern, for the CPU baseline ("port") ----------------------
def reference_style_loop(p_open, p_closed, base_params, alpha, forecast_steps=20):
    """Per-sample Python loop with scipy.odeint exactly as predict_batch step 2 does
    (06_lstm_ode_integration.py:372-401).  Serial by construction; used as cpu_baseline."""
    from scipy.integrate import odeint
    n = len(p_open)
    trajs = np.empty((n, forecast_steps, 3), dtype=np.float64)
    t = np.linspace(0, forecast_steps, forecast_steps)
    for i in range(n):
        po, pc = float(p_open[i]), float(p_closed[i])
        if pc > 0.6:
            y0 = [0.2, 0.2, 0.6]
        elif po > 0.6:
            y0 = [0.6, 0.2, 0.2]
        else:
            y0 = [0.33, 0.34, 0.33]
        prm = dict(base_params)
        prm["k_af"] *= (1 + alpha * pc)
        prm["k_pf"] *= (1 + alpha * pc)
        prm["k_fa"] *= (1 + alpha * po)
        prm["k_pa"] *= (1 + alpha * po)
        for kk in prm:
            prm[kk] = max(RATE_FLOOR, prm[kk])

        def f(y, _t, prm=prm):
            A, P, F = max(0, y[0]), max(0, y[1]), max(0, y[2])
            return [-prm["k_ap"] * A - prm["k_af"] * A + prm["k_pa"] * P + prm["k_fa"] * F,
                    prm["k_ap"] * A - prm["k_pa"] * P - prm["k_pf"] * P + prm["k_fp"] * F,
                    prm["k_af"] * A + prm["k_pf"] * P - prm["k_fa"] * F - prm["k_fp"] * F]

        y0 = np.array(y0) / np.sum(y0)
        sol = odeint(f, y0, t)
        sol = np.clip(sol, 0, 1)
        trajs[i] = sol / sol.sum(axis=1, keepdims=True)
    return trajs


# ---- time-varying rates (05_ode_model.py:171-196) ------------------------------------------
def solve_with_modulation(params, initial_state, t_span, modulation_func, n_points=100):
    """Restatement of CognitiveStateODE.solve_with_modulation: LSODA (scipy.odeint defaults) on a right-hand side that
    asks `modulation_func(t, params.copy())` for the rate dict at every evaluation (05:188-190), then clip + renormalise
    (05:193-194).  Returns (t, solution[n_points,3])."""
    from scipy.integrate import odeint
    t = np.linspace(t_span[0], t_span[1], n_points)                    # 05:183
    y0 = np.array(initial_state, dtype=np.float64) / np.sum(initial_state)   # 05:184

    def f(y, tt):
        p = modulation_func(tt, dict(params))
        k = np.array([[p[name]] for name in RATE_ORDER], dtype=np.float64)
        return rhs(np.asarray(y, dtype=np.float64).reshape(1, 3), k, True)[0]

    return t, post_process_06(odeint(f, y0, t))


def rk4_modulated(style, y0, nodes, t_span, n_points, substeps):
    """Fixed-step RK4 reading the rates from a node table sampled at t0 + m*h/2 (what bci_ode_solve_modulated
    integrates): y0 (N,3), nodes (M,6) or (M,6,N), M = 2*substeps*(n_points-1)+1.  Returns (N,n_points,3)."""
    y = np.asarray(y0, dtype=np.float64).reshape(-1, 3)
    clamp = style == STYLE_REF06
    if clamp:
        y = y / y.sum(axis=1, keepdims=True)
    nodes = np.asarray(nodes, dtype=np.float64)
    if nodes.ndim == 2:
        nodes = np.broadcast_to(nodes[:, :, None], nodes.shape + (y.shape[0],))
    assert nodes.shape[0] == 2 * substeps * (n_points - 1) + 1
    out = np.empty((y.shape[0], n_points, 3), dtype=np.float64)
    out[:, 0] = y
    h = (t_span[1] - t_span[0]) / (n_points - 1) / substeps
    m = 0
    for i in range(1, n_points):
        for _ in range(substeps):
            k1 = rhs(y, nodes[m], clamp)
            k2 = rhs(y + 0.5 * h * k1, nodes[m + 1], clamp)
            k3 = rhs(y + 0.5 * h * k2, nodes[m + 1], clamp)
            k4 = rhs(y + h * k3, nodes[m + 2], clamp)
            y = y + (h / 6.0) * ((k1 + k4) + 2.0 * (k2 + k3))
            m += 2
        out[:, i] = y
    return post_process_06(out) if clamp else out


def modulation_cases():
    """Named modulation functions shared by tests/golden/make_golden_modulation.py (run against the live reference) and the
    parity tests: (name, modulation_func, initial_state, t_span, n_points)."""
    def lstm_coupling(t, p):        # time-varying P(closed) driving the 06-style coupling (06:249-262), alpha = 0.5
        pc = 0.5 + 0.4 * np.sin(0.35 * t)
        po = 1.0 - pc
        p["k_af"] *= 1 + 0.5 * pc; p["k_pf"] *= 1 + 0.5 * pc
        p["k_fa"] *= 1 + 0.5 * po; p["k_pa"] *= 1 + 0.5 * po
        return {k: max(RATE_FLOOR, v) for k, v in p.items()}

    def fatigue_ramp(t, p):         # time-on-task: fatigue inflow grows, recovery decays
        p["k_af"] *= 1.0 + 0.08 * t
        p["k_fa"] *= np.exp(-0.05 * t)
        return p

    def identity(t, p):
        return p

    return [("lstm_coupling", lstm_coupling, [0.6, 0.2, 0.2], (0.0, 20.0), 100),
            ("fatigue_ramp", fatigue_ramp, [0.33, 0.34, 0.33], (0.0, 50.0), 60),
            ("identity", identity, [0.2, 0.2, 0.6], (5.0, 25.0), 20)]
