"""ctypes binding of ``libprego_b200.so`` (the C ABI declared in ``include/prego_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C prego_b200/csrc``.
There is no fallback: if the shared object is missing or a call fails, a ``RuntimeError``
is raised -- the product path never routes through a CPU implementation.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libprego_b200.so")

PREC_BF16 = 0
PREC_FP32 = 1
PREC_F16 = 2
PREC_TF32 = 3  # training only
PACK_F32, PACK_16, PACK_X3, PACK_ALL = 1, 2, 4, 7  # PREGO_PACK_* (include/prego_b200.h)
PREC_F16X3 = 4  # fp32-class accuracy on the tensor cores (split fp16 operands)
PRECISIONS = {"bf16": PREC_BF16, "fp32": PREC_FP32, "fp16": PREC_F16, "fp16x3": PREC_F16X3}
PREC_TF32X3 = 5  # training only: three-term TF32 split, fp32-class
TRAIN_PRECISIONS = {"fp32": PREC_FP32, "tf32": PREC_TF32, "tf32x3": PREC_TF32X3}
PHASES = ("stage", "gemm1", "layernorm", "gemm2", "recurrence", "head")


class Dims(C.Structure):
    _fields_ = [("d_rgb", C.c_int32), ("d_flow", C.c_int32), ("embed_dim", C.c_int32),
                ("hidden_dim", C.c_int32), ("num_classes", C.c_int32)]


class Weights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "layer1_0_weight", "layer1_0_bias", "layer1_1_weight", "layer1_1_bias",
        "gru_weight_ih_l0", "gru_weight_hh_l0", "gru_bias_ih_l0", "gru_bias_hh_l0",
        "f_classification_0_weight", "f_classification_0_bias")]


class ForwardArgs(C.Structure):
    _fields_ = [("rgb", C.c_void_p), ("flow", C.c_void_p), ("B", C.c_int64), ("T", C.c_int64),
                ("h_state", C.c_void_p), ("probs", C.c_void_p), ("logits", C.c_void_p),
                ("labels", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
                ("precision", C.c_int32), ("chunk_T", C.c_int32), ("feature_dtype", C.c_int32), ("flow_is_zero", C.c_int32)]


FEAT_F32, FEAT_16 = 0, 1


class AnticipationArgs(C.Structure):
    _fields_ = [("probs", C.c_void_p), ("logits", C.c_void_p), ("labels", C.c_void_p), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_size_t)]


class Grads(C.Structure):
    _fields_ = Weights._fields_


class TrainArgs(C.Structure):
    _fields_ = [("rgb", C.c_void_p), ("flow", C.c_void_p), ("B", C.c_int64), ("T", C.c_int64),
                ("logits", C.c_void_p), ("dlogits", C.c_void_p), ("grads", C.POINTER(Grads)),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("dropout_p", C.c_float),
                ("seed", C.c_uint64), ("precision", C.c_int32), ("reserved", C.c_int32), ("gru_grads_event", C.c_void_p)]


ADAMW_MAX_TENSORS = 16


class AdamWArgs(C.Structure):
    _fields_ = [("params", C.c_void_p * ADAMW_MAX_TENSORS), ("grads", C.c_void_p * ADAMW_MAX_TENSORS),
                ("exp_avg", C.c_void_p * ADAMW_MAX_TENSORS), ("exp_avg_sq", C.c_void_p * ADAMW_MAX_TENSORS),
                ("numel", C.c_int64 * ADAMW_MAX_TENSORS), ("num_tensors", C.c_int32), ("step", C.c_int32),
                ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("weight_decay", C.c_float), ("grad_scale", C.c_float)]


# name -> (restype, argtypes); every symbol include/prego_b200.h declares
SIGNATURES = {
    "prego_abi_version": (C.c_int, []),
    "prego_last_error": (C.c_char_p, []),
    "prego_model_create": (C.c_int, [C.POINTER(Dims), C.c_int32, C.POINTER(C.c_void_p)]),
    "prego_model_destroy": (C.c_int, [C.c_void_p]),
    "prego_model_load_weights": (C.c_int, [C.c_void_p, C.POINTER(Weights), C.c_void_p]),
    "prego_model_load_weights_ex": (C.c_int, [C.c_void_p, C.POINTER(Weights), C.c_uint32, C.c_void_p]),
    "prego_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32]),
    "prego_forward": (C.c_int, [C.c_void_p, C.POINTER(ForwardArgs), C.c_void_p]),
    "prego_model_load_anticipation": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "prego_anticipation_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64, C.c_int32]),
    "prego_forward_anticipation": (C.c_int, [C.c_void_p, C.POINTER(ForwardArgs), C.POINTER(AnticipationArgs), C.c_void_p]),
    "prego_ap_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "prego_perframe_ap": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "prego_host_round_features": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32]),
    "prego_host_round_impl": (C.c_int, []),
    "prego_host_all_zero": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32]),
    "prego_host_stager_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_void_p)]),
    "prego_host_stager_destroy": (C.c_int, [C.c_void_p]),
    "prego_host_stager_run": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "prego_online_open": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.POINTER(C.c_void_p)]),
    "prego_online_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "prego_online_wait": (C.c_int, [C.c_void_p]),
    "prego_online_step_wait": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "prego_online_close": (C.c_int, [C.c_void_p]),
    "prego_online_trace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "prego_device_error": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "prego_recurrence_fallbacks": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "prego_train_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64, C.c_int64]),
    "prego_train_workspace_bytes_ex": (C.c_size_t, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32]),
    "prego_train_forward": (C.c_int, [C.c_void_p, C.POINTER(TrainArgs), C.c_void_p]),
    "prego_train_backward": (C.c_int, [C.c_void_p, C.POINTER(TrainArgs), C.c_void_p]),
    "prego_adamw_step": (C.c_int, [C.POINTER(AdamWArgs), C.c_void_p]),
    "prego_profile_begin": (C.c_int, [C.c_void_p]),
    "prego_profile_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "prego_window_mode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32,
                                    C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "prego_rle": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p,
                            C.c_void_p, C.c_void_p, C.c_void_p]),
    "prego_gemm16_nt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                  C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "prego_gemm16_stats_nt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                        C.c_int64, C.c_int32, C.c_float, C.c_void_p]),
    "prego_gemm16_ln_nt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_int64, C.c_int64, C.c_int32, C.c_void_p]),
    "prego_gemm_tf32_nt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                     C.c_int32, C.c_void_p]),
    "prego_gemm_f32_nt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                    C.c_int64, C.c_void_p]),
}

_lib = None


def load():
    """Load the shared library (once) and attach prototypes.  Raises RuntimeError if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C prego_b200/csrc`. prego_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().prego_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")
