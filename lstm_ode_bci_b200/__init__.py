"""Importable alias of the package directory `lstm-ode-bci_b200/` (a hyphen cannot appear
in a Python module name).  All code lives there; this file only extends the search path."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                 "lstm-ode-bci_b200"))
from ._version import __version__  # noqa: E402,F401
