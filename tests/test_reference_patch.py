"""Container-only drop-in check (needs /root/reference; runs with -m "not gpu", skipped on the GPU box): import the reference's
own scripts, apply patch.patch_reference, and verify that (1) every hot-path name the script defines is replaced by the B200
implementation and (2) each replacement accepts the reference's call signature -- same parameter names, same defaults -- so the
script's unmodified main()/pipeline code keeps calling it the way it does today.  No GPU work happens here."""
import inspect

import pytest

from oracle import ref_loader

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present")]

EXPECT = {
    "ref02": {"bandpass_filter", "normalize_data", "create_sequences"},
    "ref04": {"EnhancedLSTMModel"},
    "ref05": {"CognitiveStateODE", "sensitivity_analysis"},
    "ref06": {"EnhancedLSTMModel", "CognitiveStateODE", "LSTMODEIntegration"},
    "ref07": {"EnhancedLSTMModel", "compute_channel_importance", "compute_permutation_importance"},
    "ref08": {"EnhancedLSTMModel", "multistep_forecast", "rolling_forecast_evaluation", "prob_to_ode_state",
              "predict_trajectory", "get_lstm_probabilities"},
    "ref09": {"AblationLSTMModel"},
    "ref10": {"EnhancedLSTMModel", "CognitiveStateODE", "get_three_state_probabilities"},
}

# methods whose call contract the pipeline scripts rely on (SURVEY.md section 8 b)
METHODS = {
    "EnhancedLSTMModel": ["__init__", "forward"],
    "AblationLSTMModel": ["__init__", "forward"],
    "CognitiveStateODE": ["__init__", "ode_system", "solve"],
    "LSTMODEIntegration": ["__init__", "get_lstm_probabilities", "modulate_ode_rates", "predict_trajectory", "predict_batch"],
}


def _accepts(ref_fn, new_fn, where):
    """every parameter of the reference callable exists in the replacement with the same default (extra keyword parameters
    with defaults are allowed)"""
    rp, np_ = inspect.signature(ref_fn).parameters, inspect.signature(new_fn).parameters
    for name, p in rp.items():
        if p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD):
            continue
        assert name in np_, "%s: parameter %r of the reference is missing" % (where, name)
        # (a parameter the reference requires may be optional here: every reference call still binds the same way)
        assert np_[name].default == p.default or p.default is inspect._empty, \
            "%s: default of %r is %r, reference has %r" % (where, name, np_[name].default, p.default)
    ref_pos = [n for n, p in rp.items() if p.kind == p.POSITIONAL_OR_KEYWORD]
    new_pos = [n for n, p in np_.items() if p.kind == p.POSITIONAL_OR_KEYWORD]
    assert new_pos[:len(ref_pos)] == ref_pos, "%s: positional order %s != reference %s" % (where, new_pos, ref_pos)
    for name, p in np_.items():
        if name not in rp and p.kind not in (p.VAR_POSITIONAL, p.VAR_KEYWORD):
            assert p.default is not inspect._empty, "%s: extra parameter %r has no default" % (where, name)


@pytest.mark.parametrize("name", sorted(EXPECT))
def test_patch_reference_replaces_hot_path_names(name):
    from lstm_ode_bci_b200 import patch
    mod = ref_loader.load(name, fresh=True)
    originals = {n: getattr(mod, n) for n in EXPECT[name] if hasattr(mod, n)}
    assert set(originals) == EXPECT[name], "reference script %s no longer defines %s" % (name, EXPECT[name] - set(originals))
    done = set(patch.patch_reference(mod))
    assert EXPECT[name] <= done, EXPECT[name] - done
    for n, ref_obj in originals.items():
        new_obj = getattr(mod, n)
        assert new_obj is not ref_obj and new_obj.__module__.startswith("lstm_ode_bci_b200"), n
        if inspect.isclass(ref_obj):
            for meth in METHODS.get(n, []):
                if hasattr(ref_obj, meth):
                    _accepts(getattr(ref_obj, meth), getattr(new_obj, meth), "%s.%s.%s" % (name, n, meth))
        else:
            _accepts(ref_obj, new_obj, "%s.%s" % (name, n))


def test_patched_model_keeps_the_reference_state_dict_abi():
    """a checkpoint written by the reference class loads into the drop-in with strict=True (04:921-933 / 06:416-430)"""
    import torch
    from lstm_ode_bci_b200 import lstm
    ref04 = ref_loader.load("ref04")
    torch.manual_seed(0)
    ref_model = ref04.EnhancedLSTMModel(61, 128, 3, 2, 0.4, True)
    ours = lstm.EnhancedLSTMModel(61, 128, 3, 2, 0.4, True)
    missing = ours.load_state_dict(ref_model.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert [k for k, _ in ours.named_parameters()] == [k for k, _ in ref_model.named_parameters()]
