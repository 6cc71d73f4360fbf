"""ctypes binding of ``liblsdm_b200.so`` (C ABI declared in ``include/lsdm_b200.h``).

The product path has NO fallback: if the CUDA library is missing or fails to load, every
entry point raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C lsdm_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblsdm_b200.so")

OK, EINVAL, ESTATE, ECUDA, ENOMEM = 0, -1, -2, -3, -4


class LsdmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"lsdm_b200 error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("batch_local", C.c_int32), ("batch_global", C.c_int32), ("batch_offset", C.c_int32),
                ("n_cats", C.c_int32), ("device", C.c_int32), ("reserved", C.c_int32)]


_P = C.c_void_p


class CfMhaWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w_q", "b_q", "w_k", "b_k", "w_v", "b_v", "fc_w", "fc_b", "ln_w", "ln_b")]


class CfFfnWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w1", "b1", "w2", "b2", "ln_w", "ln_b")] + [("d_hid", C.c_int32), ("reserved", C.c_int32)]


# name -> (restype, argtypes); exactly the prototypes of include/lsdm_b200.h
PROTOTYPES = {
    "lsdm_version": (C.c_char_p, []),
    "lsdm_last_error": (C.c_char_p, []),
    "lsdm_create": (C.c_int, [C.POINTER(_P), C.POINTER(Config)]),
    "lsdm_destroy": (None, [_P]),
    "lsdm_set_batch": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32]),
    "lsdm_load_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int32, _P]),
    "lsdm_num_weights": (C.c_int, [_P]),
    "lsdm_weight_key": (C.c_char_p, [_P, C.c_int]),
    "lsdm_finalize_weights": (C.c_int, [_P, _P]),
    "lsdm_set_schedule": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int32, _P]),
    "lsdm_workspace_bytes": (C.c_size_t, [_P]),
    "lsdm_set_workspace": (C.c_int, [_P, _P, C.c_size_t]),
    "lsdm_encode_conditions": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "lsdm_encode_conditions_train": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
    "lsdm_set_allreduce": (C.c_int, [_P, _P, _P]),
    "lsdm_read_weight": (C.c_int, [_P, C.c_char_p, _P, C.c_int64, _P]),
    "lsdm_denoise_step": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int32, _P]),
    "lsdm_forward": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "lsdm_sample_loop": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P]),
    "lsdm_get_out_cat": (C.c_int, [_P, _P, _P]),
    "lsdm_get_pcd_out": (C.c_int, [_P, _P, _P]),
    "lsdm_q_sample": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "lsdm_chamfer": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "lsdm_cat_loss": (C.c_int, [_P, _P, _P, C.c_int32, _P, _P]),
    "lsdm_train_tape_bytes": (C.c_size_t, [_P]),
    "lsdm_grad_floats": (C.c_int64, [_P]),
    "lsdm_weight_slot": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "lsdm_training_backward": (C.c_int, [_P] + [_P] * 10 + [C.c_float, C.c_float, C.c_float, _P, C.c_size_t, _P, _P, _P, _P]),
    "lsdm_adamw_step": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int64, C.c_float, _P]),
    "lsdm_eval_emd": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "lsdm_eval_fscore": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_double, _P, _P, _P]),
    "lsdm_eval_chamfer": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "lsdm_eval_topk": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, C.c_int32, _P, _P]),
    "lsdm_clip_create": (C.c_int, [C.POINTER(_P), C.c_int32]),
    "lsdm_clip_destroy": (None, [_P]),
    "lsdm_clip_load_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int32, _P]),
    "lsdm_clip_finalize": (C.c_int, [_P]),
    "lsdm_clip_dims": (C.c_int, [_P] + [C.POINTER(C.c_int32)] * 6),
    "lsdm_clip_set_precision": (C.c_int, [_P, C.c_int32]),
    "lsdm_clip_workspace_bytes": (C.c_size_t, [_P, C.c_int32, C.c_int32]),
    "lsdm_clip_encode_text": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, C.c_size_t, _P, _P]),
    "lsdm_clip_launch_count": (C.c_int64, [_P]),
    "lsdm_cf_set_option": (C.c_int, [C.c_char_p, C.c_int32]),
    "lsdm_cf_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "lsdm_cf_mha_forward": (C.c_int, [C.POINTER(CfMhaWeights), _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P,
                                      C.c_size_t, _P, _P]),
    "lsdm_cf_ffn_forward": (C.c_int, [C.POINTER(CfFfnWeights), _P, C.c_int64, C.c_int32, _P, C.c_size_t, _P, _P]),
    "lsdm_debug_tensor": (C.c_int64, [_P, C.c_char_p, _P, C.c_size_t, _P]),
    "lsdm_launch_count": (C.c_int64, [_P]),
    "lsdm_set_option": (C.c_int, [_P, C.c_char_p, C.c_int32]),
    "lsdm_set_precision": (C.c_int, [_P, C.c_int32, C.c_int32]),
    "lsdm_debug_gemm": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int64, _P, C.c_int64, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_int32, C.c_int32, C.c_int32, _P]),
    "lsdm_profile_begin": (C.c_int, [_P]),
    "lsdm_profile_report": (C.c_char_p, [_P]),
    "lsdm_profile_end": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int32, C.POINTER(C.c_double)]),
}

ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int32)

_lib = None


def load():
    """Load the shared library (once) and bind every prototype.  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LsdmError(ESTATE, f"{LIB_PATH} not built: the CUDA path is the only path (no CPU fallback). "
                                "Run `make -C lsdm_b200/csrc` or `python -c 'import __graft_entry__ as g; g.build()'`.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != OK:
        raise LsdmError(code, load().lsdm_last_error().decode())
