"""ctypes binding of libbci_b200.so -- the only way Python reaches the CUDA kernels.

There is deliberately no CPU fallback: if the shared library is missing or the device is not
a B200-class GPU (sm_100), calls raise.  Mirrors include/bci_b200.h one to one.
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
ABI_VERSION = 5   # include/bci_b200.h: BCI_ABI_VERSION
LIB_PATH = os.path.join(_PKG, "lib", "libbci_b200.so")

BCI_MAX_LAYERS = 4
PRECISION_FP32, PRECISION_BF16 = 0, 1
ODE_RK4, ODE_RK45 = 0, 1
STYLE_REF06, STYLE_REF08 = 0, 1
Y0_GIVEN, Y0_FROM_PROBS_06, Y0_FROM_PCLOSED_08 = 0, 1, 2
OUT_F32, OUT_F64 = 0, 1
IN_F32, IN_BF16 = 0, 1
COMM_HANDLE_BYTES = 128

_ERR_NAMES = {-1: "BCI_EINVAL", -2: "BCI_ECUDA", -3: "BCI_ENOMEM", -4: "BCI_ESTATE", -5: "BCI_EUNSUPPORTED"}


class BciError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (_ERR_NAMES.get(code, str(code)), msg))
        self.code = code


class LstmConfig(C.Structure):
    _fields_ = [("input_size", C.c_int32), ("hidden_size", C.c_int32), ("num_layers", C.c_int32),
                ("num_classes", C.c_int32), ("bidirectional", C.c_int32), ("precision", C.c_int32),
                ("use_attention", C.c_int32), ("use_layer_norm", C.c_int32)]


_FP = C.c_void_p  # device pointers travel as integers
_LAYER_ARR = (_FP * 2) * BCI_MAX_LAYERS

_WEIGHT_FIELDS = (
    [("input_proj_w", _FP), ("input_proj_b", _FP), ("input_ln_w", _FP), ("input_ln_b", _FP),
     ("w_ih", _LAYER_ARR), ("w_hh", _LAYER_ARR), ("b_ih", _LAYER_ARR), ("b_hh", _LAYER_ARR),
     ("ln_w", _FP), ("ln_b", _FP), ("attn_w1", _FP), ("attn_b1", _FP), ("attn_w2", _FP), ("attn_b2", _FP),
     ("cls_w0", _FP), ("cls_b0", _FP), ("cls_w3", _FP), ("cls_b3", _FP), ("cls_w6", _FP), ("cls_b6", _FP)])


class LstmWeights(C.Structure):
    _fields_ = _WEIGHT_FIELDS


class LstmGrads(C.Structure):
    _fields_ = _WEIGHT_FIELDS


class LstmInput(C.Structure):
    _fields_ = [("data", _FP), ("dtype", C.c_int32), ("windows_per_run", C.c_int32), ("window_stride", C.c_int64),
                ("run_stride", C.c_int64), ("first_window", C.c_int64)]


class OdeArgs(C.Structure):
    _fields_ = [("mode", C.c_int32), ("style", C.c_int32), ("y0_mode", C.c_int32), ("coupling", C.c_int32),
                ("n", C.c_int64), ("base_rates", C.c_float * 6), ("alpha", C.c_float),
                ("rates", _FP), ("alpha_arr", _FP), ("p_open", _FP), ("p_closed", _FP), ("y0", _FP),
                ("t_end", C.c_double), ("n_points", C.c_int32), ("substeps", C.c_int32),
                ("rtol", C.c_double), ("atol", C.c_double), ("out_dtype", C.c_int32),
                ("traj", _FP), ("final_state", _FP), ("n_steps", _FP)]


class OdeModArgs(C.Structure):
    _fields_ = [("n", C.c_int64), ("style", C.c_int32), ("n_points", C.c_int32), ("substeps", C.c_int32),
                ("per_trajectory", C.c_int32), ("t_span", C.c_double), ("rate_nodes", _FP), ("y0", _FP),
                ("traj", _FP), ("final_state", _FP)]


class PreprocArgs(C.Structure):
    _fields_ = [("n_recordings", C.c_int32), ("n_channels", C.c_int32), ("n_samples", C.c_int64), ("in_dtype", C.c_int32),
                ("order", C.c_int32), ("b_host", C.POINTER(C.c_double)), ("a_host", C.POINTER(C.c_double)),
                ("zi_host", C.POINTER(C.c_double)), ("padlen", C.c_int32), ("seq_len", C.c_int32), ("step", C.c_int32),
                ("mean_in", _FP), ("std_in", _FP)]


# name -> (restype, argtypes); every symbol include/bci_b200.h declares
SIGNATURES = {
    "bci_abi_version": (C.c_int, []),
    "bci_last_error": (C.c_char_p, []),
    "bci_device_check": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    "bci_lstm_create": (C.c_int, [C.POINTER(LstmConfig), C.POINTER(C.c_void_p)]),
    "bci_lstm_destroy": (C.c_int, [C.c_void_p]),
    "bci_lstm_load_weights": (C.c_int, [C.c_void_p, C.POINTER(LstmWeights), C.c_void_p]),
    "bci_lstm_set_profiling": (C.c_int, [C.c_void_p, C.c_int32]),
    "bci_lstm_get_profile": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "bci_launch_count": (C.c_int64, []),
    "bci_lstm_chunk_windows": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "bci_lstm_workspace_bytes": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "bci_lstm_forward": (C.c_int, [C.c_void_p, _FP, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_uint64,
                                   _FP, _FP, _FP, _FP, C.c_size_t, C.c_void_p]),
    "bci_lstm_forward_view": (C.c_int, [C.c_void_p, C.POINTER(LstmInput), C.c_int32, C.c_int32, _FP, _FP, _FP, _FP, C.c_size_t,
                                        C.c_void_p]),
    "bci_ce_loss_grad": (C.c_int, [_FP, _FP, _FP, C.c_int32, C.c_int32, C.c_float, _FP, _FP, C.c_void_p]),
    "bci_grad_accumulate": (C.c_int, [_FP, _FP, C.c_int64, C.c_int32, C.c_void_p]),
    "bci_lstm_backward": (C.c_int, [C.c_void_p, _FP, _FP, C.c_int32, C.c_int32, _FP, C.POINTER(LstmGrads),
                                    _FP, C.c_size_t, C.c_void_p]),
    "bci_adamw_step": (C.c_int, [_FP, _FP, _FP, _FP, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float,
                                 C.c_float, C.c_int32, C.c_float, C.c_float, _FP, C.c_void_p]),
    "bci_comm_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_void_p)]),
    "bci_comm_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bci_comm_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bci_comm_bucket": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "bci_comm_destroy": (C.c_int, [C.c_void_p]),
    "bci_fused_step": (C.c_int, [C.c_void_p, _FP, _FP, _FP, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                 C.c_int32, C.c_float, _FP, C.c_void_p]),
    "bci_preprocess_workspace_bytes": (C.c_int, [C.POINTER(PreprocArgs), C.POINTER(C.c_size_t)]),
    "bci_preprocess": (C.c_int, [C.POINTER(PreprocArgs), _FP, _FP, _FP, _FP, _FP, _FP, C.c_size_t, C.c_void_p]),
    "bci_ode_solve": (C.c_int, [C.POINTER(OdeArgs), C.c_void_p]),
    "bci_ode_solve_modulated": (C.c_int, [C.POINTER(OdeModArgs), C.c_void_p]),
    "bci_ode_classify": (C.c_int, [_FP, C.c_int64, _FP, _FP, C.c_void_p]),
    "bci_ode_forecast_readout": (C.c_int, [_FP, C.c_int64, C.c_int32, C.POINTER(C.c_int32), C.c_int32, _FP, C.c_void_p]),
    "bci_selftest_proj_gemm_bf16": (C.c_int, [_FP, _FP, _FP, _FP, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "bci_selftest_proj_gemm_bf16_blocked": (C.c_int, [_FP, _FP, _FP, _FP, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "bci_selftest_rec_bf16": (C.c_int, [_FP, _FP, _FP, _FP, C.c_int32, C.c_int32, C.c_void_p]),
    "bci_selftest_rec256_bf16": (C.c_int, [_FP, _FP, _FP, C.c_int32, C.c_int32, C.c_void_p]),
    "bci_selftest_fused_rec_bf16": (C.c_int, [_FP, _FP, _FP, _FP, _FP, _FP, _FP, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "bci_selftest_gemm_tf32x3": (C.c_int, [C.c_int32, _FP, _FP, _FP, _FP, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_void_p]),
    "bci_selftest_gemm_tf32_single": (C.c_int, [_FP, _FP, _FP, _FP, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "bci_selftest_gemm_tf32_single_tn": (C.c_int, [_FP, _FP, _FP, C.c_int32, C.c_int32, C.c_int64, C.c_void_p]),
    "bci_selftest_gemm_f16x3": (C.c_int, [_FP, _FP, _FP, _FP, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "bci_selftest_rec_f16x3": (C.c_int, [_FP, _FP, _FP, _FP, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "bci_selftest_rec_swap_fwd": (C.c_int, [_FP, _FP, _FP, _FP, _FP, _FP, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "bci_selftest_bptt_swap": (C.c_int, [_FP, _FP, _FP, _FP, _FP, _FP, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "bci_selftest_rec_swap256_fwd": (C.c_int, [_FP, _FP, _FP, _FP, _FP, _FP, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "bci_selftest_bptt_swap256": (C.c_int, [_FP, _FP, _FP, _FP, _FP, _FP, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "bci_selftest_dropout_mask": (C.c_int, [_FP, C.c_int64, C.c_float, C.c_uint64, C.c_uint32, C.c_void_p]),
    "bci_selftest_swap_set_debug": (C.c_int, [_FP]),
    "bci_selftest_tmem_a_probe": (C.c_int, [_FP, C.c_void_p]),
    "bci_lstm_set_train_mode": (C.c_int, [C.c_void_p, C.c_int32]),
    "bci_host_stage": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32]),
    "bci_permute_channels": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                             C.c_int32, C.c_void_p, C.c_void_p]),
    "bci_fp32_peak_probe": (C.c_int, [C.POINTER(C.c_double), C.c_void_p]),
    "bci_fp64_peak_probe": (C.c_int, [C.POINTER(C.c_double), C.c_void_p]),
}

_lib = None


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            # a fresh checkout: try the in-tree nvcc build once (this is the build step, not a fallback -- without a
            # working CUDA toolchain the import fails loudly)
            try:
                from . import build as _build
                import sys
                print("[bci_b200] libbci_b200.so missing; building it in-tree with nvcc ...", file=sys.stderr)
                _build.build()
            except Exception as e:
                raise BciError(-2, "libbci_b200.so not found at %s and the in-tree build failed (%s) -- run `python -m "
                                   "lstm_ode_bci_b200.build` (or __graft_entry__.build()); there is no CPU fallback" % (LIB_PATH, e))
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if l.bci_abi_version() != ABI_VERSION:
            raise BciError(-1, "ABI version mismatch")
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise BciError(rc, lib().bci_last_error().decode("utf-8", "replace"))


def require_device(device_index=0):
    """Raise unless `device_index` is an sm_100 GPU; returns the SM count."""
    sm = C.c_int(0)
    check(lib().bci_device_check(int(device_index), C.byref(sm)))
    return sm.value
