"""ctypes binding of libccal.so (include/ccal.h).  There is no fallback: if the library is
missing or the device is not sm_100 every product call raises."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_longlong, c_uint32, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libccal.so")

CCAL_F32, CCAL_F16, CCAL_BF16 = 0, 1, 2
FX_SHIFT = 40
MAX_THRESHOLDS = 63
MAX_K = 16

# name -> (restype, argtypes); mirrors include/ccal.h one to one
SIGNATURES = {
    "ccal_version": (c_int, []),
    "ccal_last_error": (c_char_p, []),
    "ccal_launch_count": (c_longlong, []),
    "ccal_check_device": (c_int, []),
    "ccal_trace_marks_report": (c_int, [c_char_p, c_int]),
    "ccal_score_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int64, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_double), c_int,
                                 c_void_p, c_void_p]),
    "ccal_score_guess_stats": (c_int, [POINTER(ctypes.c_ulonglong), c_int]),
    "ccal_score_trace": (c_int, [POINTER(ctypes.c_ulonglong), c_int]),
    "ccal_score_pass1": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ccal_score_pass2": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int64, c_int, c_int, c_int, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, POINTER(c_double), c_int, c_void_p, c_void_p]),
    "ccal_ts_loss_grad": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int64, c_int, c_int, c_int,
                                  c_void_p, c_void_p, c_void_p]),
    "ccal_ts_loss_grad_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int,
                                      c_void_p, c_void_p, c_void_p]),
    "ccal_sgd_scalar_step": (c_int, [c_void_p, c_void_p, c_double, c_double, c_double, c_void_p]),
    "ccal_knn_l2": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_void_p,
                            c_void_p]),
    "ccal_knn_l2_exhaustive": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_void_p,
                                       c_void_p]),
    "ccal_dac_fit": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ccal_dac_fit_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ccal_dac_predict_logits": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "ccal_logits_confidence": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "ccal_dac_softmax_logits": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "ccal_row_argmax": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "ccal_bin_stats": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int64, POINTER(c_double), c_int,
                               c_void_p, POINTER(c_double), c_int, c_void_p, c_void_p]),
    "ccal_radix_hist": (c_int, [c_void_p, c_int64, c_int, POINTER(c_uint32), c_int, c_void_p, c_void_p]),
    "ccal_class_counts": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "ccal_exp_normalise_rows": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ccal_isotonic_fit_binary": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, POINTER(c_int64), c_void_p]),
    "ccal_isotonic_transform": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_double, c_void_p, c_void_p]),
    "ccal_ova_hist_fit": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "ccal_ova_apply": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_int, c_void_p, c_void_p]),
    "ccal_sort_pairs_f64_u8": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "ccal_prefix_sum_i32": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "ccal_kde2_pdf": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_double, c_double, c_void_p,
                              c_void_p]),
    "ccal_density_ratio_apply": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_double, c_void_p,
                                         c_void_p, c_void_p, c_void_p]),
}

_lib = None


class CcalError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load libccal.so (building is a separate, explicit step: python -m clip_calibration_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CcalError(
            f"{LIB_PATH} not found. Build it with `python -m clip_calibration_b200.build` "
            "(needs nvcc); this package has no CPU or PyTorch fallback path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().ccal_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    msg = last_error()
    if rc == 1:
        raise ValueError(f"{what}: {msg}")
    raise CcalError(f"{what} failed (code {rc}): {msg}")


def doubles(values):
    n = len(values)
    arr = (c_double * max(n, 1))(*[float(v) for v in values])
    return arr, n


def uint32s(values):
    n = len(values)
    arr = (c_uint32 * max(n, 1))(*[int(v) for v in values])
    return arr, n
