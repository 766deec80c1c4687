"""patch_reference(module): swap the hot-path callables of an imported reference script
(04/06/08/10) for the B200 implementations, so its own main()/pipeline code runs unmodified.

The reference redeclares its classes in every script, so we patch by structure:
whatever the module calls `EnhancedLSTMModel`, `CognitiveStateODE`, `LSTMODEIntegration`,
`predict_trajectory`, `prob_to_ode_state`, `multistep_forecast`, `rolling_forecast_evaluation`,
`get_lstm_probabilities`, `get_three_state_probabilities` is replaced when present.
"""
from . import integration, lstm, ode

_REPLACEMENTS = {
    "EnhancedLSTMModel": lstm.EnhancedLSTMModel,
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
