"""ctypes binding of include/e2t.h (the C-ABI drop-in boundary).

The product path loads exactly one library: ``libe2t.so`` built by nvcc for sm_100a next to this
file (``python __graft_entry__.py build``).  There is no CPU fallback: if the library is missing,
or no CUDA device is present, loading / ``e2t_create`` raises.
"""
from __future__ import annotations

import ctypes as C
import os

E2T_MAX_SUBNETS = 16
E2T_MAX_LAYERS = 8
HOST, DEVICE, STAGED0, STAGED1 = 0, 1, 2, 3
VALUE, GRAD, ADAM_M, ADAM_V, EMA = 0, 1, 2, 3, 4
GRAD_AND_COUNT = 5   # e2t_flat_buffer only: gradients + [token count, 0, 0, 0] -- ONE all-reduce per data-parallel step
ACT = {"linear": 0, "relu": 1}
GEMM = {"auto": 0, "simt": 1, "tcgen05": 2}
ATTN = {"none": 0, "luong": 1, "bahdanau": 2}
AUX_KIND = {"gaussian": 0, "categorical": 1}
ABI_VERSION = 9


class E2TConfig(C.Structure):
    _fields_ = [
        ("n_subnets", C.c_int32),
        ("subnet_id", C.c_int32 * E2T_MAX_SUBNETS),
        ("subnet_C", C.c_int32 * E2T_MAX_SUBNETS),
        ("subnet_W", C.c_int32 * E2T_MAX_SUBNETS),
        ("E", C.c_int32),
        ("n_enc_layers", C.c_int32),
        ("H", C.c_int32 * E2T_MAX_LAYERS),
        ("D", C.c_int32),
        ("Hd", C.c_int32),
        ("V", C.c_int32),
        ("conv_act", C.c_int32),
        ("emb_act", C.c_int32),
        ("pad_id", C.c_int32),
        ("eos_id", C.c_int32),
        ("start_id", C.c_int32),
        ("max_B", C.c_int32),
        ("max_T", C.c_int32),
        ("max_L", C.c_int32),
        ("max_beam", C.c_int32),
        ("ff_dropout", C.c_float),
        ("rnn_dropout", C.c_float),
        ("lr", C.c_float),
        ("beta1", C.c_float),
        ("beta2", C.c_float),
        ("eps", C.c_float),
        ("ema_decay", C.c_float),
        ("penalty_scale", C.c_float),
        ("gemm_backend", C.c_int32),
        ("device", C.c_int32),
        ("attention", C.c_int32),
        ("aux_layer", C.c_int32),
        ("aux_hidden", C.c_int32),
        ("aux_F", C.c_int32),
        ("aux_kind", C.c_int32),
        ("aux_penalty", C.c_float),
        ("proj_hidden", C.c_int32),
    ]


_P = C.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "e2t_last_error": (C.c_char_p, []),
    "e2t_abi_version": (C.c_int, []),
    "e2t_create": (C.c_int, [C.POINTER(E2TConfig), C.POINTER(_P)]),
    "e2t_destroy": (C.c_int, [_P]),
    "e2t_set_stream": (C.c_int, [_P, _P]),
    "e2t_sync": (C.c_int, [_P]),
    "e2t_param_count": (C.c_int, [_P]),
    "e2t_param_total": (C.c_int64, [_P]),
    "e2t_param_info": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int),
                                 C.POINTER(C.c_int64)]),
    "e2t_get_tensor": (C.c_int, [_P, C.c_char_p, C.c_int, _P]),
    "e2t_set_tensor": (C.c_int, [_P, C.c_char_p, C.c_int, _P]),
    "e2t_flat_buffer": (C.c_int, [_P, C.c_int, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "e2t_set_trainable": (C.c_int, [_P, C.c_char_p, C.c_int]),
    "e2t_get_step": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "e2t_set_step": (C.c_int, [_P, C.c_int64]),
    "e2t_train_step_grads": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32,
                                       C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "e2t_stage_inputs": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_int]),
    "e2t_read_loss_accumulators": (C.c_int, [_P, C.POINTER(C.c_double), C.c_int]),
    "e2t_wait_staged": (C.c_int, [_P, C.c_int]),
    "e2t_host_alloc": (C.c_int, [C.POINTER(_P), C.c_int64]),
    "e2t_host_free": (C.c_int, [_P]),
    "e2t_set_grad_buckets": (C.c_int, [_P, C.c_int]),
    "e2t_grad_bucket_count": (C.c_int, [_P]),
    "e2t_grad_bucket_info": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "e2t_grad_bucket_wait": (C.c_int, [_P, C.c_int, _P]),
    "e2t_set_encoder_targets": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int]),
    "e2t_last_losses": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "e2t_post_losses": (C.c_int, [_P, C.c_int]),
    "e2t_fetch_losses": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "e2t_input_saliency": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                     C.c_float, _P, _P]),
    "e2t_adam_ema_step": (C.c_int, [_P, C.c_int, C.c_float]),
    "e2t_adam_ema_step_dev": (C.c_int, [_P, C.c_int, _P]),
    "e2t_eval_loss": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "e2t_greedy_decode": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                    _P, _P]),
    "e2t_beam_decode": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_float, _P, _P]),
    "e2t_get_activation": (C.c_int, [_P, C.c_char_p, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "e2t_launch_counts": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "e2t_counter": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_int64)]),
    "e2t_profile_enable": (C.c_int, [_P, C.c_int]),
    "e2t_profile_read": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "e2t_bench_gemm": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.POINTER(C.c_float)]),
    "e2t_profile_report": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "e2t_selftest_gemm": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libe2t.so")


def bind(lib: C.CDLL) -> C.CDLL:
    """Attach restype/argtypes for every symbol of e2t.h; raises if one is missing."""
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    ver = lib.e2t_abi_version()
    if ver != ABI_VERSION:
        raise RuntimeError(f"libe2t ABI version {ver} != expected {ABI_VERSION}; rebuild the extension")
    return lib


_lib = None


def load() -> C.CDLL:
    """The product library.  Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build the CUDA extension first "
                "(python -c 'import __graft_entry__ as g; g.build()').  There is no CPU fallback.")
        _lib = bind(C.CDLL(LIB_PATH))
    return _lib
