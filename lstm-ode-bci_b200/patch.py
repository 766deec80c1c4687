"""patch_reference(module): swap the hot-path callables of an imported reference script
(02/04/05/06/07/08/09/10) for the B200 implementations, so its own main()/pipeline code runs unmodified.

The reference redeclares its classes in every script, so we patch by structure:
whatever the module calls `EnhancedLSTMModel`, `CognitiveStateODE`, `LSTMODEIntegration`,
`predict_trajectory`, `prob_to_ode_state`, `multistep_forecast`, `rolling_forecast_evaluation`,
`get_lstm_probabilities`, `get_three_state_probabilities`, 05's `sensitivity_analysis`, 07's `compute_channel_importance` /
`compute_permutation_importance` is replaced when present.
"""
import functools
import inspect

from . import explain, integration, lstm, ode, preprocessing

_REPLACEMENTS = {
    "EnhancedLSTMModel": lstm.EnhancedLSTMModel,
    "AblationLSTMModel": lstm.AblationLSTMModel,            # 09_sensitivity_analysis.py:176-240
    "bandpass_filter": preprocessing.bandpass_filter,       # 02_preprocessing.py:114-131
    "normalize_data": preprocessing.normalize_data,         # 02:134-154
    "create_sequences": preprocessing.create_sequences,     # 02:157-180
    "CognitiveStateODE": ode.CognitiveStateODE,
    "LSTMODEIntegration": integration.LSTMODEIntegration,
    "get_three_state_probabilities": integration.get_three_state_probabilities,
    "multistep_forecast": integration.multistep_forecast,
    "rolling_forecast_evaluation": integration.rolling_forecast_evaluation,
    "prob_to_ode_state": integration.prob_to_ode_state,
    "sensitivity_analysis": ode.sensitivity_analysis,       # 05_ode_model.py:687-750 (12 steady states in one launch)
}
# 07_explainability.py:203-361: the attribution sweeps; they label channels with the script's own EEG_CHANNELS (07:63-71,222-225)
_EXPLAIN = {
    "compute_channel_importance": explain.compute_channel_importance,
    "compute_permutation_importance": explain.compute_permutation_importance,
}


def _with_reference_defaults(new_obj, ref_obj):
    """The reference redeclares its classes per script with DIFFERENT defaults (EnhancedLSTMModel: input_size 14 / hidden 128 in
    04/07/08, 64 / 256 in 06/10): the replacement installed into a script takes that script's defaults."""
    new_fn = new_obj.__init__ if inspect.isclass(new_obj) else new_obj
    ref_fn = ref_obj.__init__ if inspect.isclass(ref_obj) else ref_obj
    try:
        ref_sig, new_sig = inspect.signature(ref_fn), inspect.signature(new_fn)
    except (TypeError, ValueError):
        return new_obj
    over = {n: p.default for n, p in ref_sig.parameters.items()
            if p.default is not inspect._empty and n in new_sig.parameters and new_sig.parameters[n].default != p.default}
    if not over:
        return new_obj
    sig = new_sig.replace(parameters=[p.replace(default=over[n]) if n in over else p for n, p in new_sig.parameters.items()])

    def fill(args, kwargs, skip_self):
        given = new_sig.bind_partial(*(((None,) if skip_self else ()) + tuple(args)), **kwargs).arguments
        return dict(kwargs, **{n: v for n, v in over.items() if n not in given})

    if inspect.isclass(new_obj):
        def __init__(self, *args, **kwargs):
            new_obj.__init__(self, *args, **fill(args, kwargs, True))
        __init__.__signature__ = sig
        return type(new_obj.__name__, (new_obj,), {"__init__": __init__, "__module__": new_obj.__module__,
                                                   "__doc__": new_obj.__doc__, "__qualname__": new_obj.__qualname__})

    @functools.wraps(new_obj)
    def wrapper(*args, **kwargs):
        return new_obj(*args, **fill(args, kwargs, False))
    wrapper.__signature__ = sig
    return wrapper


def _with_channel_names(fn, names):
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        kwargs.setdefault("channel_names", names)
        return fn(*args, **kwargs)
    return wrapper


def patch_reference(module):
    """Returns the list of names replaced.  08's module-level predict_trajectory/get_lstm_probabilities
    are only replaced when the module has no LSTMODEIntegration (i.e. it is 08, not 06)."""
    done = []
    for name, repl in _REPLACEMENTS.items():
        if hasattr(module, name):
            setattr(module, name, _with_reference_defaults(repl, getattr(module, name)))
            done.append(name)
    for name, repl in _EXPLAIN.items():
        if hasattr(module, name):
            fn = _with_reference_defaults(repl, getattr(module, name))
            names = getattr(module, "EEG_CHANNELS", None)
            if names is not None:
                fn = _with_channel_names(fn, list(names))
            setattr(module, name, fn)
            done.append(name)
    if not hasattr(module, "LSTMODEIntegration"):
        for name in ("predict_trajectory", "get_lstm_probabilities"):
            if hasattr(module, name):
                setattr(module, name, _with_reference_defaults(getattr(integration, name), getattr(module, name)))
                done.append(name)
    return done
