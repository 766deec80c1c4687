"""CPU checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports
every symbol include/bci_b200.h declares; the ctypes mirror covers the same set; the product
package never imports the oracle."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from lstm_ode_bci_b200 import build
    return build.build()


def _declared():
    src = open(os.path.join(ROOT, "include", "bci_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bci_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(libpath):
    out = subprocess.run(["nm", "-D", "--defined-only", libpath], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (bci_[a-z0-9_]+)", out))
    declared = _declared()
    assert len(declared) >= 14
    assert not [s for s in declared if s not in exported]


def test_ctypes_mirror_matches_header(libpath):
    from lstm_ode_bci_b200 import _native
    assert sorted(_native.SIGNATURES) == _declared()
    lib = _native.lib()
    hdr = open(os.path.join(ROOT, "include", "bci_b200.h")).read()
    assert lib.bci_abi_version() == _native.ABI_VERSION == int(re.search(r"#define BCI_ABI_VERSION (\d+)", hdr).group(1))
    for name in _native.SIGNATURES:
        assert getattr(lib, name) is not None


def test_graft_entry_build(libpath):
    """The driver's build hook: compiles (or finds current) the library, imports the package, checks the ABI version."""
    import __graft_entry__ as g
    assert g.build() == libpath


def test_library_is_sm100a_only(libpath):
    out = subprocess.run(["cuobjdump", "-lelf", libpath], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_struct_layouts_match_c(libpath, tmp_path):
    """sizeof/offsetof of the ABI structs as the C compiler sees them == the ctypes mirror."""
    import ctypes as C
    from lstm_ode_bci_b200 import _native as N
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "bci_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                    "sizeof(bci_lstm_config),sizeof(bci_lstm_weights),sizeof(bci_lstm_grads),sizeof(bci_ode_args),"
                    "offsetof(bci_ode_args,rates),offsetof(bci_ode_args,t_end),offsetof(bci_ode_args,traj),"
                    "sizeof(bci_ode_mod_args),offsetof(bci_ode_mod_args,t_span),offsetof(bci_ode_mod_args,final_state),"
                    "sizeof(bci_preproc_args),offsetof(bci_preproc_args,std_in),sizeof(bci_lstm_input),offsetof(bci_lstm_input,window_stride));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(N.LstmConfig), C.sizeof(N.LstmWeights), C.sizeof(N.LstmGrads), C.sizeof(N.OdeArgs),
            N.OdeArgs.rates.offset, N.OdeArgs.t_end.offset, N.OdeArgs.traj.offset,
            C.sizeof(N.OdeModArgs), N.OdeModArgs.t_span.offset, N.OdeModArgs.final_state.offset,
            C.sizeof(N.PreprocArgs), N.PreprocArgs.std_in.offset, C.sizeof(N.LstmInput), N.LstmInput.window_stride.offset]
    assert got == want


def test_no_cpu_fallback_and_no_oracle_in_product():
    import torch
    from lstm_ode_bci_b200 import _native, lstm, ode
    pkg = os.path.join(ROOT, "lstm-ode-bci_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "/root/reference" not in txt, f
    m = lstm.EnhancedLSTMModel(61, 128, 3, 2)
    with pytest.raises(_native.BciError):
        m(torch.zeros(1, 8, 61))                       # CPU tensor: must raise, not fall back
    with pytest.raises(_native.BciError):
        ode.solve_ensemble(1, y0=torch.zeros(3, 1), device="cpu")


def test_state_dict_abi_matches_reference_names():
    from lstm_ode_bci_b200 import lstm, synth
    for H in (128, 256):
        m = lstm.EnhancedLSTMModel(61, H, 3, 2)
        shapes = synth.lstm_param_shapes(61, H, 3, 2)
        sd = m.state_dict()
        assert list(sd.keys()) == list(shapes.keys())
        for k, v in sd.items():
            assert tuple(v.shape) == tuple(shapes[k]), k
    assert sum(p.numel() for p in lstm.EnhancedLSTMModel(61, 128, 3, 2).parameters()) == 1137731
    assert sum(p.numel() for p in lstm.EnhancedLSTMModel(61, 256, 3, 2).parameters()) == 4520067


def test_host_scalar_helpers_match_golden(golden):
    import numpy as np
    from lstm_ode_bci_b200 import integration, ode, synth
    g = golden("ode_ref08.npz")
    y0 = np.stack([integration.prob_to_ode_state(p) for p in g["p_closed"]])
    assert np.array_equal(y0.astype(np.float64), g["y0"])
    g6 = golden("ode_ref06.npz")
    for i in range(0, 96, 7):
        base = {k: float(g6["base"][j, i]) for j, k in enumerate(synth.RATE_ORDER)}
        integ = integration.LSTMODEIntegration(None, ode.CognitiveStateODE(dict(base)), float(g6["alpha"][i]))
        mod = integ.modulate_ode_rates(g6["p_closed"][i], g6["p_open"][i])
        assert [mod[k] for k in synth.RATE_ORDER] == list(g6["rates"][:, i])


def test_train_mode_constants_match_header():
    """ops.TRAIN_MODES mirrors the BCI_TRAIN_* enum of include/bci_b200.h; bci_lstm_set_train_mode is exported."""
    from lstm_ode_bci_b200 import _native, ops
    hdr = open(os.path.join(ROOT, "include", "bci_b200.h")).read()
    m = re.search(r"enum \{ BCI_TRAIN_FP32 = (\d+), BCI_TRAIN_MIXED = (\d+) \}", hdr)
    assert m and ops.TRAIN_MODES == {"fp32": int(m.group(1)), "mixed": int(m.group(2))}
    assert hasattr(_native.lib(), "bci_lstm_set_train_mode")
