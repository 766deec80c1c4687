"""Import the reference's numbered scripts as modules (TEST INFRASTRUCTURE ONLY).

The reference (khurrameycon/LSTM-ODE-BCI, mounted read-only at /root/reference in the
build container) is a set of stand-alone scripts whose names start with a digit, so they
cannot be imported with `import`.  They also import matplotlib/seaborn at module top,
which are not installed here and are only used by plotting functions.  This loader
stubs those modules and loads a script by path.

Used only by tests/golden/make_golden.py and by container-only tests that compare the
oracle with the live reference.  /root/reference does not exist on the GPU box: nothing
under tests marked `gpu`, bench.py or smoke() may call this.
"""
import importlib.util
import io
import contextlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("BCI_REFERENCE_ROOT", "/root/reference")

_FILES = {
    "ref02": "02_preprocessing.py",
    "ref09": "09_sensitivity_analysis.py",
    "ref04": "04_lstm_model.py",
    "ref05": "05_ode_model.py",
    "ref06": "06_lstm_ode_integration.py",
    "ref07": "07_explainability.py",
    "ref08": "08_forecasting.py",
    "ref10": "10_three_state_probabilities.py",
}


def reference_available():
    return os.path.isdir(REFERENCE_ROOT) and os.path.isfile(os.path.join(REFERENCE_ROOT, _FILES["ref04"]))


def _install_plot_stubs():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.patches", "seaborn", "mne"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                m = types.ModuleType(name)
                m.__dict__.setdefault("__path__", [])
                sys.modules[name] = m
    # attribute wiring so "import matplotlib.pyplot as plt" works on the stub
    mpl = sys.modules["matplotlib"]
    for sub in ("pyplot", "gridspec", "patches"):
        if not hasattr(mpl, sub):
            setattr(mpl, sub, sys.modules["matplotlib." + sub])
    if not hasattr(mpl, "use"):
        mpl.use = lambda *a, **k: None
    plt = sys.modules["matplotlib.pyplot"]
    if not hasattr(plt, "style"):
        plt.style = types.SimpleNamespace(use=lambda *a, **k: None)
    if not hasattr(plt, "rcParams"):
        plt.rcParams = {}
    gs = sys.modules["matplotlib.gridspec"]
    if not hasattr(gs, "GridSpec"):                 # 07_explainability.py:33 `from matplotlib.gridspec import GridSpec`
        gs.GridSpec = type("GridSpec", (), {})
    mne = sys.modules["mne"]
    if not hasattr(mne, "set_log_level"):
        mne.set_log_level = lambda *a, **k: None
    sns = sys.modules["seaborn"]
    for fn in ("set_style", "set_palette", "set_theme", "set_context"):
        if not hasattr(sns, fn):
            setattr(sns, fn, lambda *a, **k: None)


_cache = {}


def load(name, fresh=False):
    """name in {ref02, ref04, ref05, ref06, ref07, ref08, ref09, ref10}; returns the imported module (fresh=True: a new, uncached
    module object -- for tests that monkey-patch it)."""
    if name in _cache and not fresh:
        return _cache[name]
    if not reference_available():
        raise FileNotFoundError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_plot_stubs()
    path = os.path.join(REFERENCE_ROOT, _FILES[name])
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    # The scripts mkdir <parent-of-reference>/outputs/... at import time; suppress that
    # side effect (we never write outside the repo) together with their banner prints.
    import pathlib
    real_mkdir = pathlib.Path.mkdir
    pathlib.Path.mkdir = lambda self, *a, **k: None
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            spec.loader.exec_module(mod)
    finally:
        pathlib.Path.mkdir = real_mkdir
    if not fresh:
        _cache[name] = mod
    return mod
