"""ctypes mirror of include/hvi_b200.h and loader of the CUDA shared library.

The library is the product: if it is missing or cannot be loaded this module raises --
there is no CPU fallback on this path.
"""
from __future__ import annotations

import ctypes as C
import os

HVI_MAX_PAR = 8
HVI_MAX_LAYERS = 8

HVI_OK, HVI_EINVAL, HVI_ECUDA, HVI_ENONFINITE, HVI_ENODATA, HVI_ENOMEM, HVI_ENCCL = range(7)
HVI_F32, HVI_F64 = 0, 1
HVI_MEM_HOST, HVI_MEM_DEVICE = 0, 1
HVI_TR_IDENTITY, HVI_TR_EXP, HVI_TR_LOGISTIC = 0, 1, 2
HVI_PR_NONE, HVI_PR_NORMAL, HVI_PR_LOGNORMAL = 0, 1, 2
HVI_SRC_P, HVI_SRC_M, HVI_SRC_FIX = 0, 1, 2
HVI_ACT_IDENTITY, HVI_ACT_TANH, HVI_ACT_LOGISTIC = 0, 1, 2
HVI_SCALE_NORMAL, HVI_SCALE_RANGE = 0, 1
HVI_FLAG_OMIT_PRIORS = 1
HVI_APPROX_MEAN, HVI_APPROX_MEANVAR_SEP = 0, 1


class hvi_layer(C.Structure):
    _fields_ = [("n_in", C.c_int32), ("n_out", C.c_int32), ("act", C.c_int32), ("has_bias", C.c_int32)]


class hvi_prior(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("a", C.c_double), ("b", C.c_double)]


class hvi_desc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("dtype", C.c_int32), ("device", C.c_int32), ("flags", C.c_int32),
        ("n_P", C.c_int32), ("n_M", C.c_int32), ("n_obs", C.c_int32), ("n_covar", C.c_int32),
        ("n_layers", C.c_int32),
        ("layers", hvi_layer * HVI_MAX_LAYERS),
        ("scale_kind", C.c_int32), ("_pad0", C.c_int32),
        ("scale_a", C.c_double * HVI_MAX_PAR), ("scale_b", C.c_double * HVI_MAX_PAR),
        ("trans_P", C.c_int32 * HVI_MAX_PAR), ("trans_M", C.c_int32 * HVI_MAX_PAR),
        ("prior_P", hvi_prior * HVI_MAX_PAR), ("prior_M", hvi_prior * HVI_MAX_PAR),
        ("n_cor_ends_P", C.c_int32), ("cor_ends_P", C.c_int32 * HVI_MAX_PAR),
        ("n_cor_ends_M", C.c_int32), ("cor_ends_M", C.c_int32 * HVI_MAX_PAR),
        ("slot_src", C.c_int32 * 4), ("slot_idx", C.c_int32 * 4), ("slot_fix", C.c_double * 4),
        ("n_pbm_covar", C.c_int32), ("pbm_covar_idx", C.c_int32 * HVI_MAX_PAR),
        ("count_globals", C.c_int32),
        ("approx", C.c_int32),
        ("n_site_total", C.c_int64),
    ]


PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "csrc", "libhvi_b200.so")

_lib = None

# every symbol include/hvi_b200.h declares: (name, restype, argtypes)
_P, _I32, _I64, _U64, _INT = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_int
_DP = C.POINTER(C.c_double)
SYMBOLS = [
    ("hvi_version", _INT, []),
    ("hvi_last_error", C.c_char_p, [_P]),
    ("hvi_plan_create", _INT, [C.POINTER(hvi_desc), C.POINTER(_P)]),
    ("hvi_plan_destroy", _INT, [_P]),
    ("hvi_phi_len", _I64, [_P]),
    ("hvi_phi_positions", _INT, [_P, C.POINTER(_I64), C.POINTER(_I64)]),
    ("hvi_vec2utri_pos", _INT, [_I64, C.POINTER(_I64), C.POINTER(_I64)]),
    ("hvi_utri2vec_pos", _I64, [_I64, _I64]),
    ("hvi_cor_count", _I64, [C.POINTER(_I32), _I32]),
    ("hvi_plan_set_data", _INT, [_P, _I64, _P, _P, _P, _P, _INT]),
    ("hvi_plan_set_data_bcast", _INT, [_P, _I64, _P, _P, _INT, _P, _P, _INT, _INT]),
    ("hvi_plan_prefetch_data", _INT, [_P, _I64, _P, _P, _P, _P]),
    ("hvi_plan_commit_data", _INT, [_P]),
    ("hvi_plan_set_sites", _INT, [_P, C.POINTER(_I64), _I64]),
    ("hvi_plan_set_dataset", _INT, [_P, _I64, _P, _P, _P, _P, _INT]),
    ("hvi_plan_select_range", _INT, [_P, _I64, _I64]),
    ("hvi_plan_select_batch", _INT, [_P, _P, _I64, _INT]),
    ("hvi_plan_set_stream", _INT, [_P, _P, _INT]),
    ("hvi_plan_set_pbm", _INT, [_P, C.c_char_p, _I32, _I32, _DP]),
    ("hvi_pbm_compile_check", _INT, [C.c_char_p, _I32, _I32, _I32, _I32, _I32, C.c_char_p, _I32]),
    ("hvi_neg_elbo_grad", _INT, [_P, _I32, _P, _P, _P, _INT, _DP, _DP, _P]),
    ("hvi_neg_elbo_grad_rng", _INT, [_P, _I32, _P, _U64, _I64, _INT, _DP, _DP, _P]),
    ("hvi_neg_elbo_grad_async", _INT, [_P, _I32, _P, _P, _P, _U64, _I64, _P, _P]),
    ("hvi_nccl_unique_id", _INT, [_P]),
    ("hvi_plan_comm_init", _INT, [_P, _P, _I32, _I32]),
    ("hvi_plan_comm_destroy", _INT, [_P]),
    ("hvi_allreduce_result", _INT, [_P, _P, _P]),
    ("hvi_neg_elbo_grad_sharded", _INT, [_P, _I32, _P, _P, _P, _U64, _I64, _INT, _DP, _DP, _P]),
    ("hvi_loss_gf_grad", _INT, [_P, _P, _INT, _DP, _P]),
    ("hvi_adam_init", _INT, [_P, _P, C.c_double, C.c_double, C.c_double, C.c_double]),
    ("hvi_adam_run", _INT, [_P, _I32, _I32, _U64, _I64, _DP]),
    ("hvi_adam_apply_dev", _INT, [_P, _P, _P]),
    ("hvi_adam_run_minibatches", _INT, [_P, _I32, _I32, _U64, _I64, _I64, _DP]),
    ("hvi_plan_use_graphs", _INT, [_P, _INT]),
    ("hvi_plan_use_fused", _INT, [_P, _INT]),
    ("hvi_plan_fused_phase_clocks", _INT, [_P, C.POINTER(_I64)]),
    ("hvi_adam_phi_dev", _INT, [_P, C.POINTER(_P)]),
    ("hvi_adam_get_phi", _INT, [_P, _P]),
    ("hvi_generate_zeta", _INT, [_P, _I32, _P, _P, _P, _INT, _P, _P, _P]),
    ("hvi_predict", _INT, [_P, _I32, _P, _P, _P, _U64, _I64, _INT, _P, _P, _P, _DP]),
    ("hvi_apply_g", _INT, [_P, _P, _INT, _P]),
    ("hvi_randn", _INT, [_P, _I32, _I64, _U64, _I64, _INT, _P]),
    ("hvi_dense_tc", _INT, [_I32, _I64, _I32, _I32, _P, _P, _P, _I32, _P]),
    ("hvi_gemm_bf3", _INT, [_I32, _I32, _I64, _I32, _I64, _P, _P, _P, _I32, _P, _P]),
    ("hvi_launch_count", _I64, [_P]),
    ("hvi_plan_set_profiling", _INT, [_P, _INT]),
    ("hvi_plan_kernel_times", _INT, [_P, C.POINTER(C.c_float)]),
]


def load_library(path: str | None = None):
    """dlopen the C-ABI library and bind the prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} is not built -- run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the neg-ELBO path)")
    lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.hvi_version() != 1:
        raise RuntimeError("hvi_b200 ABI version mismatch")
    if path is None:
        _lib = lib
    return lib
