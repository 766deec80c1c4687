"""patch_reference(module): swap the hot-path callables of an imported reference script
(02/04/06/08/09/10) for the B200 implementations, so its own main()/pipeline code runs unmodified.

The reference redeclares its classes in every script, so we patch by structure:
whatever the module calls `EnhancedLSTMModel`, `CognitiveStateODE`, `LSTMODEIntegration`,
`predict_trajectory`, `prob_to_ode_state`, `multistep_forecast`, `rolling_forecast_evaluation`,
`get_lstm_probabilities`, `get_three_state_probabilities` is replaced when present.
"""
from . import integration, lstm, ode, preprocessing

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
}


def patch_reference(module):
    """Returns the list of names replaced.  08's module-level predict_trajectory/get_lstm_probabilities
    are only replaced when the module has no LSTMODEIntegration (i.e. it is 08, not 06)."""
    done = []
    for name, repl in _REPLACEMENTS.items():
        if hasattr(module, name):
            setattr(module, name, repl)
            done.append(name)
    if not hasattr(module, "LSTMODEIntegration"):
        for name in ("predict_trajectory", "get_lstm_probabilities"):
            if hasattr(module, name):
                setattr(module, name, getattr(integration, name))
                done.append(name)
    return done
