"""GPU parity of the callers: predict_batch (06:308-406), predict_trajectory (06:266-306),
get_three_state_probabilities (10:204-290) against the live-reference golden vectors."""
import numpy as np
import pytest
import torch

from lstm_ode_bci_b200 import integration, lstm, ode, patch, synth

pytestmark = pytest.mark.gpu


def test_predict_batch_matches_reference(golden):
    g = golden("pipeline_h128.npz")
    gl = golden("lstm_h128.npz")
    params = synth.make_lstm_params(int(gl["seed_w"]), 61, 128, 3, logit_gain=float(gl["gain"]))
    x = synth.make_windows(int(gl["seed_x"]), int(gl["B"]), 256, 61, structured=bool(gl["structured"]))
    m = lstm.from_params(params, precision="fp32")
    integ = integration.LSTMODEIntegration(m, ode.CognitiveStateODE(), coupling_strength=0.5)
    traj, probs, preds = integ.predict_batch(x, forecast_steps=20, batch_size=3, show_progress=False)
    assert traj.shape == (8, 20, 3) and traj.dtype == np.float64
    assert probs.shape == (8, 2) and probs.dtype == np.float32 and preds.shape == (8,)
    assert np.abs(probs - g["probs"]).max() <= 1e-5
    assert np.abs(traj - g["traj"]).max() <= 1e-6
    assert np.array_equal(preds, g["preds"])
    assert integ.ode_model.params == integ.base_params          # params restored (06:404)
    tr1, pr1, at1 = integ.predict_trajectory(x[:1], forecast_steps=10)
    assert np.abs(tr1 - g["single_traj"]).max() <= 1e-6
    assert np.abs(pr1 - g["single_probs"]).max() <= 1e-5
    assert np.abs(at1 - g["single_attn"]).max() <= 1e-6
    lp, three, cls = integration.get_three_state_probabilities(m, ode.CognitiveStateODE(), x, batch_size=4)
    assert np.abs(lp - g["lstm_probs10"]).max() <= 1e-5
    assert np.abs(three - g["three_state"]).max() <= 1e-6
    assert np.array_equal(cls, g["cls"])
    # device-resident variant agrees with the host-facing one
    trd, prd, fin, pred, cls_d = integ.predict_batch_device(torch.from_numpy(x).cuda())
    assert np.abs(trd.cpu().numpy() - traj).max() <= 1e-6 and np.array_equal(pred.cpu().numpy(), preds)


def test_patch_reference_swaps_by_name():
    import types
    fake = types.SimpleNamespace(EnhancedLSTMModel=object, CognitiveStateODE=object, predict_trajectory=object,
                                 prob_to_ode_state=object, multistep_forecast=object, get_lstm_probabilities=object)
    done = patch.patch_reference(fake)
    assert fake.EnhancedLSTMModel is lstm.EnhancedLSTMModel and fake.CognitiveStateODE is ode.CognitiveStateODE
    assert fake.predict_trajectory is integration.predict_trajectory and "multistep_forecast" in done


def test_config5_pipeline_single_rank_matches_unsharded_mirrors():
    """Config 5 (08-style forecast after 06-style coupling) through parallel.forecast_pipeline_sharded on one rank ==
    the plain mirrors run back to back."""
    from lstm_ode_bci_b200 import parallel
    params = synth.make_lstm_params(3, 61, 128, 3, logit_gain=20.0)
    x = synth.make_windows(5, 70, 128, 61, structured=True)
    m = lstm.from_params(params, precision="fp32")
    integ = integration.LSTMODEIntegration(m, ode.CognitiveStateODE(), coupling_strength=0.5)
    xd = torch.from_numpy(x).cuda()
    res = parallel.forecast_pipeline_sharded(integ, xd, len(x), horizons=(5, 10, 20))
    traj, probs, preds = integ.predict_batch(x, forecast_steps=20, batch_size=32, show_progress=False)
    assert np.abs(res["probs"].cpu().numpy() - probs).max() <= 1e-6
    assert np.abs(res["traj"].cpu().numpy() - traj).max() <= 1e-6
    fc = integration.multistep_forecast(probs, integ.base_params, horizons=[5, 10, 20])
    got = res["forecast"].cpu().numpy()
    b, e = res["forecast_range"]
    assert (b, e) == (0, len(x) - 20)
    for j, h in enumerate((5, 10, 20)):
        assert np.abs(got[:, j] - fc[h]["predictions"]).max() <= 1e-6


def test_empty_and_single_window_inputs():
    """Edge cases of the caller contracts: N = 0 (the reference's loops simply do not run, 06:339,372) and N = 1."""
    import numpy as np
    import torch
    from lstm_ode_bci_b200 import integration, lstm, ode, synth
    params = synth.make_lstm_params(42, 61, 128, 3, logit_gain=12.0)
    for precision in ("fp32", "bf16"):
        m = lstm.from_params(params, precision=precision)
        integ = integration.LSTMODEIntegration(m, ode.CognitiveStateODE(), coupling_strength=0.5)
        traj, probs, preds = integ.predict_batch(np.zeros((0, 256, 61), dtype=np.float32), forecast_steps=20, show_progress=False)
        assert traj.shape == (0, 20, 3) and probs.shape == (0, 2) and preds.shape == (0,)
        with torch.no_grad():
            lg, at = m(torch.zeros((0, 256, 61), device="cuda"), return_attention=True)
        assert lg.shape == (0, 2) and at.shape == (0, 256)
        x1 = synth.make_windows(3, 1, 256, 61, structured=True)
        traj, probs, preds = integ.predict_batch(x1, forecast_steps=20, show_progress=False)
        assert traj.shape == (1, 20, 3) and abs(float(probs.sum()) - 1.0) < 1e-5 and preds.shape == (1,)
        assert np.abs(traj.sum(axis=2) - 1).max() <= 2e-7
    lp, ts, cls = integration.get_three_state_probabilities(m, ode.CognitiveStateODE(), np.zeros((0, 256, 61), dtype=np.float32))
    assert lp.shape == (0, 2) and ts.shape == (0, 3) and cls.shape == (0,)
    t, f = ode.solve_modulated_ensemble(np.zeros((3, 0)), np.tile(np.array([0.1, 0.02, 0.15, 0.08, 0.05, 0.1]), (2 * 2 * 4 + 1, 1)),
                                        (0.0, 4.0), 5, 2)
    assert t.shape == (0, 5, 3) and f.shape == (0, 3)


def test_auto_precision_follows_the_reference_autocast():
    """06_lstm_ode_integration.py:348-351 enters autocast inside predict_batch; 06:216-234, 08:203-208, 10:226-231 do not."""
    m = lstm.EnhancedLSTMModel(61, 128, 3, 2, precision="auto")
    assert m._precision_now() == "fp32"
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert m._precision_now() == "bf16"
    assert lstm.EnhancedLSTMModel(61, 128, 3, 2, precision="fp32")._precision_now() == "fp32"
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert lstm.EnhancedLSTMModel(61, 128, 3, 2, precision="fp32")._precision_now() == "fp32"     # explicit wins
        assert lstm.AblationLSTMModel(61, 256, 1, bidirectional=False, precision="auto")._precision_now() == "fp32"


def test_predict_batch_takes_the_bf16_engine_where_the_reference_autocasts():
    """precision="auto": LSTMODEIntegration.predict_batch == the explicit bf16 model bit for bit (the reference's autocast site,
    06:348-351), get_lstm_probabilities == the explicit fp32 model (no autocast there, 06:216-234)."""
    params = synth.make_lstm_params(42, 61, 128, 3, logit_gain=12.0)
    x = synth.make_windows(7, 40, 256, 61, structured=True)
    outs = {}
    for prec in ("auto", "bf16", "fp32"):
        m = lstm.from_params(params, precision=prec)
        integ = integration.LSTMODEIntegration(m, ode.CognitiveStateODE(), coupling_strength=0.5)
        outs[prec] = (integ.predict_batch(x, forecast_steps=20, batch_size=512, show_progress=False)[1], integ.get_lstm_probabilities(x)[0])
    assert np.array_equal(outs["auto"][0], outs["bf16"][0]) and not np.array_equal(outs["auto"][0], outs["fp32"][0])
    assert np.array_equal(outs["auto"][1], outs["fp32"][1])
    assert np.abs(outs["auto"][0] - outs["fp32"][0]).max() <= 1.2e-3


def test_pageable_windows_staged_as_bf16_change_no_bit():
    """Drop-in callers hand pageable fp32 numpy windows (06:346).  With the bf16 engine the native staging copy (`bci_host_stage`)
    narrows them to bf16 on the host -- the rounding the input projection applies on load anyway: probabilities and attention are
    bit-identical to the fp32-staged path and to the device-resident call; several passes, a ragged tail, attention on and off."""
    import os
    params = synth.make_lstm_params(42, 61, 128, 3, logit_gain=12.0)
    m = lstm.from_params(params, precision="bf16")
    x = synth.make_windows(11, 75, 256, 61, structured=True)
    with torch.no_grad():
        want_p, want_a = m.predict_proba(torch.from_numpy(x).cuda(), return_attention=True)
    batches = [x[:32], x[32:64], x[64:]]
    for flag in ("1", "0"):
        os.environ["BCI_STAGING_BF16"] = flag
        try:
            outs = list(integration.stream_lstm_probs(m, batches, chunk=20, want_attn=True))
            outs2 = list(integration.stream_lstm_probs(m, batches, chunk=20, want_attn=False))
        finally:
            os.environ.pop("BCI_STAGING_BF16", None)
        assert torch.equal(torch.cat([o[0] for o in outs]), want_p), flag
        assert torch.equal(torch.cat([o[1] for o in outs]), want_a), flag
        assert torch.equal(torch.cat([o[0] for o in outs2]), want_p) and all(o[1] is None for o in outs2)
    # the fp32 engine never sees narrowed input
    mf = lstm.from_params(params, precision="fp32")
    with torch.no_grad():
        want_f = mf.predict_proba(torch.from_numpy(x).cuda())
    assert torch.equal(torch.cat([o[0] for o in integration.stream_lstm_probs(mf, batches, chunk=20)]), want_f)


def test_stream_recordings_equals_materialised_windows():
    """integration.stream_recordings (host recordings -> H2D ring -> windows cut in place) == predict_proba on the windows
    create_sequences (02:157-180) would materialise; pinned and pageable host batches, fp32 and bf16 storage, fp32 and bf16 engines."""
    params = synth.make_lstm_params(42, 61, 128, 3, logit_gain=12.0)
    R, S, T, step = 3, 1504, 256, 128      # S % 8 == 0: every window start of the bf16 copy is 16-byte aligned (TMA-fed projection)
    n_seq = (S - T) // step + 1
    rng = np.random.default_rng(3)
    recs = [rng.standard_normal((R, S, 61), dtype=np.float32) for _ in range(3)]
    for prec in ("bf16", "fp32"):
        m = lstm.from_params(params, precision=prec)
        want = []
        for rec in recs:
            X = np.stack([rec[r, i * step:i * step + T] for r in range(R) for i in range(n_seq)])
            with torch.no_grad():
                want.append(m.predict_proba(torch.from_numpy(X).cuda()))
        want = torch.cat(want)
        batches = [torch.from_numpy(recs[0]).pin_memory(), recs[1], torch.from_numpy(recs[2])]      # pinned, numpy, pageable tensor
        got = torch.cat([p for p, _ in integration.stream_recordings(m, batches, seq_len=T, step=step)])
        assert got.shape == (3 * R * n_seq, 2) and torch.equal(got, want), prec
        if prec == "bf16":       # bf16 storage changes no bit in the bf16 engine
            got16 = torch.cat([p for p, _ in integration.stream_recordings(m, [torch.from_numpy(r).to(torch.bfloat16) for r in recs], T, step)])
            assert torch.equal(got16, want)
    assert list(integration.stream_recordings(m, [])) == []


def test_stream_raw_recordings_equals_preprocess_then_predict():
    """integration.stream_raw_recordings (H2D of the next batch beside filtfilt + z-score + windowing + BiLSTM of this one) == the
    same stages run one after the other; config 5 from host recordings through parallel.forecast_pipeline_from_recordings == the
    mirrors run back to back on the same windows."""
    from lstm_ode_bci_b200 import parallel, preprocessing as pp
    params = synth.make_lstm_params(3, 61, 128, 3, logit_gain=20.0)
    m = lstm.from_params(params, precision="fp32")
    rng = np.random.default_rng(5)
    raws = [(rng.standard_normal((2, 61, 4000)) * 1e-5 + 1e-4).astype(np.float32) for _ in range(4)]
    b, a, zi, padlen = pp.design_bandpass()
    want, Xs = [], []
    for raw in raws:
        out = pp.preprocess_recordings(torch.from_numpy(raw).cuda(), b, a, zi, padlen)
        Xs.append(out["X"])
        with torch.no_grad():
            want.append(m.predict_proba(out["X"]))
    want = torch.cat(want)
    got = torch.cat([p for p, _ in integration.stream_raw_recordings(m, raws)])
    assert torch.equal(got, want)
    integ = integration.LSTMODEIntegration(m, ode.CognitiveStateODE(), coupling_strength=0.5)
    n_total = want.shape[0]
    res = parallel.forecast_pipeline_from_recordings(integ, raws, n_total, raw=True)
    assert torch.equal(res["probs"], want)
    traj, probs, preds = integ.predict_batch(torch.cat(Xs).cpu().numpy(), forecast_steps=20, batch_size=32, show_progress=False)
    assert np.abs(res["traj"].cpu().numpy() - traj).max() <= 1e-6 and np.array_equal(res["pred"].cpu().numpy(), preds)
    fc = integration.multistep_forecast(probs, integ.base_params, horizons=[5, 10, 20])
    for j, h in enumerate((5, 10, 20)):
        assert np.abs(res["forecast"].cpu().numpy()[:, j] - fc[h]["predictions"]).max() <= 1e-6
