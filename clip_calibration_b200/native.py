"""Tensor-level entry points over the C ABI (include/ccal.h).

PyTorch is plumbing only: it owns device memory and the CUDA stream.  Every function takes CUDA
tensors, passes raw device pointers to libccal.so and launches on torch's current stream.
Nothing here computes on the CPU or through ATen kernels, and nothing falls back.
"""
from __future__ import annotations

import os

from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import CCAL_BF16, CCAL_F16

_DTYPES = {torch.bfloat16: CCAL_BF16, torch.float16: CCAL_F16, torch.float32: 0}
def launch_count() -> int:
    """Kernels launched by libccal.so in this process (counted inside the library at each launch site)."""
    return int(_lib.load().ccal_launch_count())


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(name: str, t: torch.Tensor, dtype=None, ndim=None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise _lib.CcalError(f"{name}: tensor is on {t.device}; this library only runs on an sm_100 GPU "
                             "(there is no CPU path)")
    if dtype is not None and t.dtype not in (dtype if isinstance(dtype, (tuple, list)) else (dtype,)):
        raise ValueError(f"{name}: dtype {t.dtype} not supported here (want {dtype})")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name}: expected {ndim} dimensions, got shape {tuple(t.shape)}")
    return t.contiguous()


def operand_dtype_for(t: torch.Tensor, requested=None):
    """Operand dtype of the fused kernels for a feature tensor.  fp16 / bf16 features go to the tensor
    cores as they are (exact products, fp32 accumulation).  fp32 (or wider) features default to
    torch.float32 = the split-precision mode: each operand becomes an fp16 pair hi + lo and every K step
    issues hi.hi + hi.lo + lo.hi, which reproduces the reference's fp32 logits to ~1e-4 absolute at s=100
    (confidences to <1e-4 relative) at 3x the tensor work.  Pass torch.float16 / torch.bfloat16 to round
    the features instead (3.7x faster; confidences then move by ~3e-3 / ~3e-2 relative)."""
    if requested is not None:
        return requested
    return t.dtype if t.dtype in (torch.float16, torch.bfloat16) else torch.float32


def new_table(n_thr: int, n_thr2: int = 0, device=None) -> torch.Tensor:
    """Zeroed bin table [(n_thr2+1), (n_thr+1), 3] (int64 view of the unsigned counters)."""
    shape = (n_thr + 1, 3) if n_thr2 == 0 else (n_thr2 + 1, n_thr + 1, 3)
    return torch.zeros(shape, dtype=torch.int64, device=device or torch.device("cuda"))


def table_to_numpy(table: torch.Tensor) -> np.ndarray:
    return table.detach().cpu().numpy().view(np.uint64)


# --------------------------------------------------------------------------------------
# K2  fused scoring
# --------------------------------------------------------------------------------------
def score_fused(img: torch.Tensor, txt: torch.Tensor, class_conf: Optional[torch.Tensor] = None,
                logit_scale: float = 100.0, labels: Optional[torch.Tensor] = None,
                thresholds: Optional[Sequence[float]] = None, table: Optional[torch.Tensor] = None,
                want_pred: bool = True, want_conf: bool = True, want_rowmax: bool = False):
    """pred / confidence (and optionally the accumulated bin table) of
    softmax(cc[pred] * logit_scale * img @ txt.T) without materialising logits.
    Returns (pred int32 [N] | None, conf float32 [N] | None, rowmax float32 [N] | None)."""
    lib = _lib.load()
    img = _need_cuda("img", img, (torch.bfloat16, torch.float16, torch.float32), 2)
    txt = _need_cuda("txt", txt, img.dtype, 2)
    if img.shape[1] != txt.shape[1]:
        raise ValueError(f"feature widths differ: img {tuple(img.shape)} vs txt {tuple(txt.shape)}")
    n, d = img.shape
    c = txt.shape[0]
    if class_conf is not None:
        class_conf = _need_cuda("class_conf", class_conf, torch.float32, 1)
        if class_conf.numel() != c:
            raise ValueError(f"class_conf has {class_conf.numel()} entries for {c} classes")
    dev = img.device
    pred = torch.empty(n, dtype=torch.int32, device=dev) if want_pred else None
    conf = torch.empty(n, dtype=torch.float32, device=dev) if want_conf else None
    rowmax = torch.empty(n, dtype=torch.float32, device=dev) if want_rowmax else None
    thr_arr, n_thr = _lib.doubles(thresholds if thresholds is not None else [])
    if table is not None:
        if labels is None or thresholds is None:
            raise ValueError("labels and thresholds are required when a bin table is requested")
        labels = _need_cuda("labels", labels, torch.int64, 1)
        table = _need_cuda("table", table, torch.int64)
        if table.numel() != 3 * (n_thr + 1):
            raise ValueError(f"table has {table.numel()} entries, expected {3 * (n_thr + 1)}")
        if labels.numel() != n:
            raise ValueError("labels length differs from the number of images")
    if n == 0:
        return pred, conf, rowmax
    with torch.cuda.device(dev):
        rc = lib.ccal_score_fused(_ptr(img), _ptr(txt), _ptr(class_conf), float(logit_scale), n, c, d,
                                  _DTYPES[img.dtype], _ptr(pred), _ptr(conf), _ptr(rowmax),
                                  _ptr(labels) if table is not None else None, thr_arr, n_thr,
                                  _ptr(table), _stream())
    _lib.check(rc, "ccal_score_fused")
    return pred, conf, rowmax


def score_guess_stats(reset: bool = False):
    """(rows scored through the FP8-guess -> bf16-verify -> redo pipeline, rows that had to be redone) on the
    current device since the last reset.  Synchronises the device."""
    import ctypes
    out = (ctypes.c_ulonglong * 2)()
    _lib.check(_lib.load().ccal_score_guess_stats(out, int(bool(reset))), "ccal_score_guess_stats")
    return int(out[0]), int(out[1])


TRACE_KINDS = ("guess_fp8", "verify_bf16", "redo", "two_pass", "temperature_scaling")


def score_trace(reset: bool = False) -> dict:
    """In-kernel timing of the scoring kernels since the last reset: {kind: {launches, ms_per_launch (first CTA in ->
    last CTA out), sm_mhz (the CTAs' cycles / their busy time: the clock the kernel really ran at)}}."""
    import ctypes
    out = (ctypes.c_ulonglong * 20)()
    _lib.check(_lib.load().ccal_score_trace(out, int(bool(reset))), "ccal_score_trace")
    res = {}
    for k, name in enumerate(TRACE_KINDS):
        launches, span, busy, cycles = (int(out[4 * k + j]) for j in range(4))
        if launches:
            res[name] = {"launches": launches, "ms_per_launch": span / launches * 1e-6,
                         "sm_mhz": (cycles / busy * 1e3) if busy else None}
    return res


def guess_pipeline_applies(n: int, c: int, d: int, dtype) -> bool:
    """Would ccal_score_fused take the FP8-guess -> bf16-verify -> redo pipeline for this shard (csrc/score_fused.cu,
    guess_pipeline_applies)?  Callers that could split a call into ccal_score_pass1 / ccal_score_pass2 (two bf16
    passes) use this to keep large shards on the cheaper one-call path."""
    env = os.environ.get("CCAL_SCORE_FP8", "")
    if dtype not in (torch.bfloat16, torch.float16) or d % 128 != 0 or d > 768 or c < 512 or env[:1] == "0":
        return False
    if env[:1] == "1":
        return True
    return n >= 128 * 2 * (torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count // 2) \
        and float(n) * float(c) >= 1.0e9


def score_pass1(img: torch.Tensor, txt: torch.Tensor):
    """First half of score_fused: (max of the RAW dot products float32 [N], first argmax int32 [N]).  Needs no
    multipliers, so it can run while the DAC fit is still in flight.  fp16 / bf16 operands."""
    lib = _lib.load()
    img = _need_cuda("img", img, (torch.bfloat16, torch.float16), 2)
    txt = _need_cuda("txt", txt, img.dtype, 2)
    if img.shape[1] != txt.shape[1]:
        raise ValueError(f"feature widths differ: img {tuple(img.shape)} vs txt {tuple(txt.shape)}")
    n, d = img.shape
    dotmax = torch.empty(n, dtype=torch.float32, device=img.device)
    pred = torch.empty(n, dtype=torch.int32, device=img.device)
    if n:
        with torch.cuda.device(img.device):
            rc = lib.ccal_score_pass1(_ptr(img), _ptr(txt), n, txt.shape[0], d, _DTYPES[img.dtype], _ptr(dotmax),
                                      _ptr(pred), _stream())
        _lib.check(rc, "ccal_score_pass1")
    return dotmax, pred


def score_pass2(img: torch.Tensor, txt: torch.Tensor, dotmax: torch.Tensor, pred: torch.Tensor,
                class_conf: Optional[torch.Tensor] = None, logit_scale: float = 100.0,
                labels: Optional[torch.Tensor] = None, thresholds: Optional[Sequence[float]] = None,
                table: Optional[torch.Tensor] = None, want_conf: bool = True):
    """Second half of score_fused from the outputs of score_pass1: confidence float32 [N] (or None) and the
    accumulated bin table; bit-identical to the one-launch form."""
    lib = _lib.load()
    img = _need_cuda("img", img, (torch.bfloat16, torch.float16), 2)
    txt = _need_cuda("txt", txt, img.dtype, 2)
    n, d = img.shape
    c = txt.shape[0]
    dotmax = _need_cuda("dotmax", dotmax, torch.float32, 1)
    pred = _need_cuda("pred", pred, torch.int32, 1)
    if dotmax.numel() != n or pred.numel() != n:
        raise ValueError("dotmax / pred must hold one entry per image")
    if class_conf is not None:
        class_conf = _need_cuda("class_conf", class_conf, torch.float32, 1)
        if class_conf.numel() != c:
            raise ValueError(f"class_conf has {class_conf.numel()} entries for {c} classes")
    conf = torch.empty(n, dtype=torch.float32, device=img.device) if want_conf else None
    thr_arr, n_thr = _lib.doubles(thresholds if thresholds is not None else [])
    if table is not None:
        if labels is None or thresholds is None:
            raise ValueError("labels and thresholds are required when a bin table is requested")
        labels = _need_cuda("labels", labels, torch.int64, 1)
        table = _need_cuda("table", table, torch.int64)
        if table.numel() != 3 * (n_thr + 1) or labels.numel() != n:
            raise ValueError("table / labels size mismatch")
    if n:
        with torch.cuda.device(img.device):
            rc = lib.ccal_score_pass2(_ptr(img), _ptr(txt), _ptr(class_conf), float(logit_scale), n, c, d,
                                      _DTYPES[img.dtype], _ptr(dotmax), _ptr(pred), _ptr(conf), None,
                                      _ptr(labels) if table is not None else None, thr_arr, n_thr, _ptr(table),
                                      _stream())
        _lib.check(rc, "ccal_score_pass2")
    return conf


# --------------------------------------------------------------------------------------
# K5  temperature-scaling loss and gradient
# --------------------------------------------------------------------------------------
def ts_loss_grad(img: torch.Tensor, txt: torch.Tensor, labels: torch.Tensor, log_scale: float):
    """(loss, d loss / d log_scale) of cross_entropy(exp(log_scale) * img @ txt.T, labels) as a
    float64 CUDA tensor of 2 elements (no host sync)."""
    lib = _lib.load()
    img = _need_cuda("img", img, (torch.bfloat16, torch.float16, torch.float32), 2)
    txt = _need_cuda("txt", txt, img.dtype, 2)
    labels = _need_cuda("labels", labels, torch.int64, 1)
    n, d = img.shape
    c = txt.shape[0]
    if labels.numel() != n:
        raise ValueError("labels length differs from the number of images")
    ws = torch.empty(2 * n, dtype=torch.float32, device=img.device)
    out = torch.empty(2, dtype=torch.float64, device=img.device)
    with torch.cuda.device(img.device):
        rc = lib.ccal_ts_loss_grad(_ptr(img), _ptr(txt), _ptr(labels), float(log_scale), n, c, d,
                                   _DTYPES[img.dtype], _ptr(ws), _ptr(out), _stream())
    _lib.check(rc, "ccal_ts_loss_grad")
    return out


def ts_sgd_step(img: torch.Tensor, txt: torch.Tensor, labels: torch.Tensor, state: torch.Tensor, lr: float,
                momentum: float, weight_decay: float) -> None:
    """One momentum-SGD step of the log-scale on the DEVICE: state (float64 CUDA [4] = {t, velocity, loss sum,
    batches}) is read by the loss / gradient kernel and updated in place; nothing returns to the host."""
    lib = _lib.load()
    img = _need_cuda("img", img, (torch.bfloat16, torch.float16), 2)
    txt = _need_cuda("txt", txt, img.dtype, 2)
    labels = _need_cuda("labels", labels, torch.int64, 1)
    if not (state.is_cuda and state.dtype == torch.float64 and state.numel() == 4 and state.is_contiguous()):
        raise ValueError("state must be a contiguous float64 CUDA tensor of 4 elements")
    n, d = img.shape
    if labels.numel() != n:
        raise ValueError("labels length differs from the number of images")
    ws = torch.empty(2 * n, dtype=torch.float32, device=img.device)
    out = torch.empty(2, dtype=torch.float64, device=img.device)
    with torch.cuda.device(img.device):
        rc = lib.ccal_ts_loss_grad_dev(_ptr(img), _ptr(txt), _ptr(labels), _ptr(state), n, txt.shape[0], d,
                                       _DTYPES[img.dtype], _ptr(ws), _ptr(out), _stream())
        _lib.check(rc, "ccal_ts_loss_grad_dev")
        rc = lib.ccal_sgd_scalar_step(_ptr(state), _ptr(out), float(lr), float(momentum), float(weight_decay), _stream())
    _lib.check(rc, "ccal_sgd_scalar_step")


# --------------------------------------------------------------------------------------
# K1  kNN + DAC fit
# --------------------------------------------------------------------------------------
def knn_l2(ref: torch.Tensor, query: torch.Tensor, k: int, drop_first: bool = False, exhaustive: bool = False):
    """k nearest `ref` rows of every `query` row: (dist [nq,k] ascending, idx [nq,k]).  Large problems
    run a tcgen05 GEMM filter + exact fp32 verification; `exhaustive=True` forces the plain exact scan."""
    lib = _lib.load()
    ref = _need_cuda("ref", ref, torch.float32, 2)
    query = _need_cuda("query", query, torch.float32, 2)
    if ref.shape[1] != query.shape[1]:
        raise ValueError("feature widths differ")
    nq, d = query.shape
    dist = torch.empty((nq, k), dtype=torch.float32, device=ref.device)
    idx = torch.empty((nq, k), dtype=torch.int32, device=ref.device)
    with torch.cuda.device(ref.device):
        fn = lib.ccal_knn_l2_exhaustive if exhaustive else lib.ccal_knn_l2
        rc = fn(_ptr(ref), _ptr(query), ref.shape[0], nq, d, int(k), int(bool(drop_first)),
                _ptr(dist), _ptr(idx), _stream())
    _lib.check(rc, "ccal_knn_l2")
    return dist, idx


def dac_fit(base_zs, cur_zs, base_tuned, cur_tuned, k: int):
    """-> (class_conf [C] f32, knn_idx_zs [C,k] i32, knn_idx_tuned, knn_dist_zs [C,k] f32, knn_dist_tuned)"""
    lib = _lib.load()
    base_zs = _need_cuda("base_text_features_zs", base_zs, torch.float32, 2)
    cur_zs = _need_cuda("current_text_features_zs", cur_zs, torch.float32, 2)
    base_tuned = _need_cuda("base_text_features_tuned", base_tuned, torch.float32, 2)
    cur_tuned = _need_cuda("current_text_features_tuned", cur_tuned, torch.float32, 2)
    b, d = base_zs.shape
    c = cur_zs.shape[0]
    if base_tuned.shape != base_zs.shape or cur_tuned.shape != cur_zs.shape or cur_zs.shape[1] != d:
        raise ValueError("zero-shot and tuned feature matrices must have matching shapes")
    dev = base_zs.device
    cc = torch.empty(c, dtype=torch.float32, device=dev)
    iz = torch.empty((c, k), dtype=torch.int32, device=dev)
    it = torch.empty((c, k), dtype=torch.int32, device=dev)
    dz = torch.empty((c, k), dtype=torch.float32, device=dev)
    dt = torch.empty((c, k), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.ccal_dac_fit(_ptr(base_zs), _ptr(cur_zs), _ptr(base_tuned), _ptr(cur_tuned), b, c, d, int(k),
                              _ptr(cc), _ptr(iz), _ptr(it), _ptr(dz), _ptr(dt), _stream())
    _lib.check(rc, "ccal_dac_fit")
    return cc, iz, it, dz, dt


def dac_fit_f16(base_zs, cur_zs, base_tuned, cur_tuned, k: int):
    """dac_fit in the reference's float16 arithmetic (inputs: float16 CUDA matrices); same five outputs, the values
    being half-precision numbers widened to float32."""
    lib = _lib.load()
    base_zs = _need_cuda("base_text_features_zs", base_zs, torch.float16, 2)
    cur_zs = _need_cuda("current_text_features_zs", cur_zs, torch.float16, 2)
    base_tuned = _need_cuda("base_text_features_tuned", base_tuned, torch.float16, 2)
    cur_tuned = _need_cuda("current_text_features_tuned", cur_tuned, torch.float16, 2)
    b, d = base_zs.shape
    c = cur_zs.shape[0]
    if base_tuned.shape != base_zs.shape or cur_tuned.shape != cur_zs.shape or cur_zs.shape[1] != d:
        raise ValueError("zero-shot and tuned feature matrices must have matching shapes")
    dev = base_zs.device
    cc = torch.empty(c, dtype=torch.float32, device=dev)
    iz = torch.empty((c, k), dtype=torch.int32, device=dev)
    it = torch.empty((c, k), dtype=torch.int32, device=dev)
    dz = torch.empty((c, k), dtype=torch.float32, device=dev)
    dt = torch.empty((c, k), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.ccal_dac_fit_f16(_ptr(base_zs), _ptr(cur_zs), _ptr(base_tuned), _ptr(cur_tuned), b, c, d, int(k),
                                  _ptr(cc), _ptr(iz), _ptr(it), _ptr(dz), _ptr(dt), _stream())
    _lib.check(rc, "ccal_dac_fit_f16")
    return cc, iz, it, dz, dt


# --------------------------------------------------------------------------------------
# K4  materialised logits
# --------------------------------------------------------------------------------------
def dac_predict_logits_(logits: torch.Tensor, class_conf: torch.Tensor) -> torch.Tensor:
    """In place: logits[i,:] *= class_conf[argmax_j logits[i,j]].  Returns pred (int32)."""
    lib = _lib.load()
    if not (logits.is_cuda and logits.is_contiguous() and logits.dtype == torch.float32 and logits.dim() == 2):
        raise ValueError("logits must be a contiguous float32 CUDA matrix (it is modified in place)")
    class_conf = _need_cuda("class_conf", class_conf, torch.float32, 1)
    n, c = logits.shape
    if class_conf.numel() != c:
        raise ValueError(f"class_conf has {class_conf.numel()} entries for {c} classes")
    pred = torch.empty(n, dtype=torch.int32, device=logits.device)
    with torch.cuda.device(logits.device):
        rc = lib.ccal_dac_predict_logits(_ptr(logits), _ptr(class_conf), n, c, _ptr(pred), _stream())
    _lib.check(rc, "ccal_dac_predict_logits")
    return pred


def logits_confidence(logits: torch.Tensor, class_conf: Optional[torch.Tensor] = None):
    lib = _lib.load()
    logits = _need_cuda("logits", logits, torch.float32, 2)
    if class_conf is not None:
        class_conf = _need_cuda("class_conf", class_conf, torch.float32, 1)
    n, c = logits.shape
    pred = torch.empty(n, dtype=torch.int32, device=logits.device)
    conf = torch.empty(n, dtype=torch.float32, device=logits.device)
    with torch.cuda.device(logits.device):
        rc = lib.ccal_logits_confidence(_ptr(logits), _ptr(class_conf), n, c, _ptr(pred), _ptr(conf), _stream())
    _lib.check(rc, "ccal_logits_confidence")
    return pred, conf


def dac_softmax_logits_(logits: torch.Tensor, class_conf: Optional[torch.Tensor] = None):
    """In place: logits[i,:] <- softmax(class_conf[pred_i] * logits[i,:]).  Returns (pred, conf)."""
    lib = _lib.load()
    if not (logits.is_cuda and logits.is_contiguous() and logits.dtype == torch.float32 and logits.dim() == 2):
        raise ValueError("logits must be a contiguous float32 CUDA matrix (it is modified in place)")
    if class_conf is not None:
        class_conf = _need_cuda("class_conf", class_conf, torch.float32, 1)
    n, c = logits.shape
    pred = torch.empty(n, dtype=torch.int32, device=logits.device)
    conf = torch.empty(n, dtype=torch.float32, device=logits.device)
    with torch.cuda.device(logits.device):
        rc = lib.ccal_dac_softmax_logits(_ptr(logits), _ptr(class_conf), n, c, _ptr(pred), _ptr(conf), _stream())
    _lib.check(rc, "ccal_dac_softmax_logits")
    return pred, conf


def row_argmax(values: torch.Tensor):
    """(first argmax int32 [N], row max float32 [N]) of an [N, C] float32 matrix."""
    lib = _lib.load()
    values = _need_cuda("values", values, torch.float32, 2)
    n, c = values.shape
    pred = torch.empty(n, dtype=torch.int32, device=values.device)
    mx = torch.empty(n, dtype=torch.float32, device=values.device)
    with torch.cuda.device(values.device):
        rc = lib.ccal_row_argmax(_ptr(values), n, c, _ptr(pred), _ptr(mx), _stream())
    _lib.check(rc, "ccal_row_argmax")
    return pred, mx


# --------------------------------------------------------------------------------------
# f-4  density-ratio calibration (2-D Gaussian KDE + row recalibration)
# --------------------------------------------------------------------------------------
def kde2_pdf(data_x: torch.Tensor, data_y: torch.Tensor, query_x: torch.Tensor, query_y: torch.Tensor,
             bw_x: float, bw_y: float) -> torch.Tensor:
    """Product-Gaussian kernel density of the float64 points (data_x, data_y) [M] at (query_x, query_y) [N]."""
    lib = _lib.load()
    data_x = _need_cuda("data_x", data_x, torch.float64, 1)
    data_y = _need_cuda("data_y", data_y, torch.float64, 1)
    query_x = _need_cuda("query_x", query_x, torch.float64, 1)
    query_y = _need_cuda("query_y", query_y, torch.float64, 1)
    if data_x.shape != data_y.shape or query_x.shape != query_y.shape:
        raise ValueError("kde2_pdf: x / y length mismatch")
    out = torch.empty(query_x.shape[0], dtype=torch.float64, device=query_x.device)
    with torch.cuda.device(query_x.device):
        rc = lib.ccal_kde2_pdf(_ptr(data_x), _ptr(data_y), data_x.shape[0], _ptr(query_x), _ptr(query_y),
                               query_x.shape[0], float(bw_x), float(bw_y), _ptr(out), _stream())
    _lib.check(rc, "ccal_kde2_pdf")
    return out


def density_ratio_apply(probs: torch.Tensor, pdf_true: torch.Tensor, pdf_false: torch.Tensor, ratio: float):
    """(probs_out float64 [N, C], conf_cal float64 [N], pred int32 [N]) - see ccal_density_ratio_apply."""
    lib = _lib.load()
    probs = _need_cuda("probs", probs, (torch.float32, torch.float64), 2)
    n, c = probs.shape
    pdf_true = _need_cuda("pdf_true", pdf_true, torch.float64, 1)
    pdf_false = _need_cuda("pdf_false", pdf_false, torch.float64, 1)
    if pdf_true.shape[0] != n or pdf_false.shape[0] != n:
        raise ValueError("density_ratio_apply: one density value per row is required")
    out = torch.empty((n, c), dtype=torch.float64, device=probs.device)
    cal = torch.empty(n, dtype=torch.float64, device=probs.device)
    pred = torch.empty(n, dtype=torch.int32, device=probs.device)
    f32 = probs.dtype == torch.float32
    with torch.cuda.device(probs.device):
        rc = lib.ccal_density_ratio_apply(_ptr(probs) if f32 else None, None if f32 else _ptr(probs), n, c,
                                          _ptr(pdf_true), _ptr(pdf_false), float(ratio), _ptr(out), _ptr(cal),
                                          _ptr(pred), _stream())
    _lib.check(rc, "ccal_density_ratio_apply")
    return out, cal, pred


# --------------------------------------------------------------------------------------
# f-4  isotonic-regression calibrator
# --------------------------------------------------------------------------------------
def exp_normalise_rows(values: torch.Tensor, labels: Optional[torch.Tensor] = None):
    """float64 exp(v) / sum(exp(v)) per row (no max shift) of an [N, C] fp32/fp64 matrix; with labels [N] also the
    one-hot matrix (uint8 [N, C])."""
    lib = _lib.load()
    values = _need_cuda("values", values, (torch.float32, torch.float64), 2)
    n, c = values.shape
    out = torch.empty((n, c), dtype=torch.float64, device=values.device)
    onehot = None
    if labels is not None:
        labels = _need_cuda("labels", labels, torch.int64, 1)
        if labels.shape[0] != n:
            raise ValueError("exp_normalise_rows: one label per row is required")
        onehot = torch.empty((n, c), dtype=torch.uint8, device=values.device)
    f32 = values.dtype == torch.float32
    with torch.cuda.device(values.device):
        rc = lib.ccal_exp_normalise_rows(_ptr(values) if f32 else None, None if f32 else _ptr(values), n, c, _ptr(out),
                                         _ptr(labels), _ptr(onehot), _stream())
    _lib.check(rc, "ccal_exp_normalise_rows")
    return out, onehot


def isotonic_fit_binary(x: torch.Tensor, y: torch.Tensor):
    """scikit-learn's isotonic fit of 0/1 targets y (uint8) at x (float64), both flat CUDA tensors ->
    (X_thresholds_, y_thresholds_) as float64 CUDA tensors.  Synchronises."""
    import ctypes
    lib = _lib.load()
    x = _need_cuda("x", x.reshape(-1), torch.float64, 1)
    y = _need_cuda("y", y.reshape(-1), torch.uint8, 1)
    if x.shape != y.shape or x.numel() == 0:
        raise ValueError("isotonic_fit_binary: x and y must be equally long and non-empty")
    n = x.numel()
    kx = torch.empty(n, dtype=torch.float64, device=x.device)
    ky = torch.empty(n, dtype=torch.float64, device=x.device)
    nk = ctypes.c_int64(0)
    with torch.cuda.device(x.device):
        rc = lib.ccal_isotonic_fit_binary(_ptr(x), _ptr(y), n, _ptr(kx), _ptr(ky), ctypes.byref(nk), _stream())
    _lib.check(rc, "ccal_isotonic_fit_binary")
    return kx[: nk.value].clone(), ky[: nk.value].clone()


def isotonic_transform(knots_x: torch.Tensor, knots_y: torch.Tensor, t: torch.Tensor,
                       residual_scale: float = 0.0) -> torch.Tensor:
    """f(clip(t)) + residual_scale * t, f = linear interpolation between the knots; same shape as t (float64)."""
    lib = _lib.load()
    knots_x = _need_cuda("knots_x", knots_x, torch.float64, 1)
    knots_y = _need_cuda("knots_y", knots_y, torch.float64, 1)
    if knots_x.shape != knots_y.shape or knots_x.numel() == 0:
        raise ValueError("isotonic_transform: knots_x / knots_y must be equally long and non-empty")
    t = _need_cuda("t", t, torch.float64)
    out = torch.empty_like(t)
    with torch.cuda.device(t.device):
        rc = lib.ccal_isotonic_transform(_ptr(knots_x), _ptr(knots_y), knots_x.numel(), _ptr(t), t.numel(),
                                         float(residual_scale), _ptr(out), _stream())
    _lib.check(rc, "ccal_isotonic_transform")
    return out


def ova_hist_fit(probs: torch.Tensor, labels: torch.Tensor, edges: torch.Tensor):
    """Per class j and bin b of `edges` (float64 [n_bins + 1] CUDA): how many rows have probs[i, j] in the bin, and how
    many of those carry the label j -> (count, hits) uint32-valued int64 CUDA tensors [C, n_bins]."""
    lib = _lib.load()
    probs = _need_cuda("probs", probs, (torch.float32, torch.float64), 2)
    n, c = probs.shape
    labels = _need_cuda("labels", labels, torch.int64, 1)
    edges = _need_cuda("edges", edges, torch.float64, 1)
    if labels.shape[0] != n:
        raise ValueError("ova_hist_fit: one label per row is required")
    n_bins = edges.numel() - 1
    both = torch.empty((2, c, n_bins), dtype=torch.int32, device=probs.device)
    f32 = probs.dtype == torch.float32
    with torch.cuda.device(probs.device):
        rc = lib.ccal_ova_hist_fit(_ptr(probs) if f32 else None, None if f32 else _ptr(probs), n, c, _ptr(labels),
                                   _ptr(edges), n_bins, _ptr(both[0]), _ptr(both[1]), _stream())
    _lib.check(rc, "ccal_ova_hist_fit")
    both = both.to(torch.int64) & 0xFFFFFFFF
    return both[0], both[1]


def ova_apply(probs: torch.Tensor, *, edges: Optional[torch.Tensor] = None, bin_map: Optional[torch.Tensor] = None,
              knots_x: Optional[torch.Tensor] = None, knots_y: Optional[torch.Tensor] = None,
              knot_off: Optional[torch.Tensor] = None, normalise: bool = True) -> torch.Tensor:
    """Column j of probs [N, C] through class j's binary calibrator - a bin map (edges + bin_map [C, n_bins]) or an
    isotonic function (concatenated knots + knot_off [C + 1] int32) - then each row divided by its sum; float64."""
    lib = _lib.load()
    probs = _need_cuda("probs", probs, (torch.float32, torch.float64), 2)
    n, c = probs.shape
    out = torch.empty((n, c), dtype=torch.float64, device=probs.device)
    n_bins = 0
    if bin_map is not None:
        edges = _need_cuda("edges", edges, torch.float64, 1)
        bin_map = _need_cuda("bin_map", bin_map, torch.float64, 2)
        n_bins = edges.numel() - 1
        if tuple(bin_map.shape) != (c, n_bins):
            raise ValueError("ova_apply: bin_map must be [C, n_bins]")
    else:
        knots_x = _need_cuda("knots_x", knots_x, torch.float64, 1)
        knots_y = _need_cuda("knots_y", knots_y, torch.float64, 1)
        knot_off = _need_cuda("knot_off", knot_off, torch.int32, 1)
        if knot_off.numel() != c + 1 or knots_x.shape != knots_y.shape:
            raise ValueError("ova_apply: knot_off must hold C + 1 offsets into equally long knots_x / knots_y")
    f32 = probs.dtype == torch.float32
    with torch.cuda.device(probs.device):
        rc = lib.ccal_ova_apply(_ptr(probs) if f32 else None, None if f32 else _ptr(probs), n, c, _ptr(edges), n_bins,
                                _ptr(bin_map), _ptr(knots_x), _ptr(knots_y), _ptr(knot_off), int(bool(normalise)),
                                _ptr(out), _stream())
    _lib.check(rc, "ccal_ova_apply")
    return out


def sort_pairs_f64_u8(keys: torch.Tensor, vals: torch.Tensor):
    """Stable ascending sort of (float64 key, uint8 payload) pairs by the library's own radix sort (the sort of the
    isotonic fit).  Synchronises."""
    lib = _lib.load()
    keys = _need_cuda("keys", keys.reshape(-1), torch.float64, 1)
    vals = _need_cuda("vals", vals.reshape(-1), torch.uint8, 1)
    if keys.shape != vals.shape:
        raise ValueError("sort_pairs_f64_u8: one payload byte per key is required")
    ko, vo = torch.empty_like(keys), torch.empty_like(vals)
    with torch.cuda.device(keys.device):
        rc = lib.ccal_sort_pairs_f64_u8(_ptr(keys), _ptr(vals), keys.numel(), _ptr(ko), _ptr(vo), _stream())
    _lib.check(rc, "ccal_sort_pairs_f64_u8")
    return ko, vo


def prefix_sum_i32(x: torch.Tensor, inclusive: bool = True) -> torch.Tensor:
    """int32 prefix sum of a flat CUDA tensor by the library's own scan kernels."""
    lib = _lib.load()
    x = _need_cuda("x", x.reshape(-1), torch.int32, 1)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = lib.ccal_prefix_sum_i32(_ptr(x), _ptr(out), x.numel(), int(bool(inclusive)), _stream())
    _lib.check(rc, "ccal_prefix_sum_i32")
    return out


# --------------------------------------------------------------------------------------
# K3  bin statistics and exact order statistics
# --------------------------------------------------------------------------------------
def bin_stats(conf: torch.Tensor, pred: torch.Tensor, gt: torch.Tensor, thresholds: Sequence[float],
              key2: Optional[torch.Tensor] = None, thresholds2: Optional[Sequence[float]] = None,
              table: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Accumulate {count, n_correct, sum conf*2^40} per bin; bin = #(thresholds <= conf)."""
    lib = _lib.load()
    conf = _need_cuda("conf", conf, (torch.float32, torch.float64), 1)
    pred = _need_cuda("pred", pred, (torch.int32, torch.int64), 1)
    gt = _need_cuda("gt", gt, torch.int64, 1)
    n = conf.numel()
    if pred.numel() != n or gt.numel() != n:
        raise ValueError("conf, pred and gt must have the same length")
    thr, n_thr = _lib.doubles(thresholds)
    thr2, n_thr2 = _lib.doubles(thresholds2 if thresholds2 is not None else [])
    if (key2 is None) != (n_thr2 == 0):
        raise ValueError("key2 and thresholds2 must be given together")
    if key2 is not None:
        key2 = _need_cuda("key2", key2, torch.float32, 1)
        if key2.numel() != n:
            raise ValueError("key2 length differs")
    if table is None:
        table = new_table(n_thr, n_thr2, conf.device)
    elif table.numel() != 3 * (n_thr + 1) * (n_thr2 + 1) or table.dtype != torch.int64 or not table.is_cuda:
        raise ValueError("table has the wrong size / dtype / device")
    with torch.cuda.device(conf.device):
        rc = lib.ccal_bin_stats(_ptr(conf), int(conf.dtype == torch.float64), _ptr(pred),
                                int(pred.dtype == torch.int64), _ptr(gt), n, thr, n_thr,
                                _ptr(key2), thr2, n_thr2, _ptr(table), _stream())
    _lib.check(rc, "ccal_bin_stats")
    return table


def class_counts(pred: torch.Tensor, gt: torch.Tensor, n_classes: int,
                 counts: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Accumulate per-class {tp, fp, fn} into an int64 [n_classes, 3] table (created zeroed if not given)."""
    lib = _lib.load()
    pred = _need_cuda("pred", pred, (torch.int32, torch.int64), 1)
    gt = _need_cuda("gt", gt, torch.int64, 1)
    if pred.shape != gt.shape:
        raise ValueError("class_counts: pred / gt length mismatch")
    if counts is None:
        counts = torch.zeros((int(n_classes), 3), dtype=torch.int64, device=pred.device)
    elif not (counts.is_cuda and counts.is_contiguous() and counts.dtype == torch.int64
              and tuple(counts.shape) == (int(n_classes), 3)):
        raise ValueError("class_counts: counts must be a contiguous int64 CUDA tensor [n_classes, 3]")
    with torch.cuda.device(pred.device):
        rc = lib.ccal_class_counts(_ptr(pred), int(pred.dtype == torch.int64), _ptr(gt), pred.shape[0],
                                   int(n_classes), _ptr(counts), _stream())
    _lib.check(rc, "ccal_class_counts")
    return counts


def radix_hist(keys: torch.Tensor, level: int, prefixes: Optional[Sequence[int]] = None,
               hist: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    keys = _need_cuda("keys", keys, torch.float32, 1)
    pf, n_pf = _lib.uint32s(prefixes if prefixes is not None else [])
    rows = 1 if level == 0 else n_pf
    if hist is None:
        hist = torch.zeros((rows, 65536), dtype=torch.int32, device=keys.device)
    with torch.cuda.device(keys.device):
        rc = lib.ccal_radix_hist(_ptr(keys), keys.numel(), int(level), pf, n_pf, _ptr(hist), _stream())
    _lib.check(rc, "ccal_radix_hist")
    return hist


def _order_statistics_bits(keys: torch.Tensor, ranks: Sequence[int], group=None):
    """(bit patterns uint32 [len(ranks)], number of keys strictly below each of them) for the given 0-based ranks of
    the globally sorted 32-bit keys (compared as unsigned bit patterns; `keys` is any 4-byte CUDA tensor viewed as
    float32).  Two 16-bit radix-histogram passes on the device, all-reduced over `group`."""
    h0 = radix_hist(keys, 0)
    if group is not None:
        torch.distributed.all_reduce(h0, group=group)
    c0 = np.cumsum(h0.cpu().numpy().astype(np.int64).ravel())
    total = int(c0[-1])
    ranks = [int(r) for r in ranks]
    if any(r < 0 or r >= total for r in ranks):
        raise ValueError("rank out of range")
    hi = np.searchsorted(c0, np.asarray(ranks), side="right")          # bucket holding each rank
    uniq = sorted(set(int(h) for h in hi))
    bits = np.empty(len(ranks), np.uint32)
    below = np.empty(len(ranks), np.int64)
    for lo in range(0, len(uniq), 64):
        part = uniq[lo:lo + 64]
        h1 = radix_hist(keys, 1, part)
        if group is not None:
            torch.distributed.all_reduce(h1, group=group)
        c1 = np.cumsum(h1.cpu().numpy().astype(np.int64), axis=1)
        for j, r in enumerate(ranks):
            if int(hi[j]) in part:
                row = part.index(int(hi[j]))
                before = int(c0[hi[j] - 1]) if hi[j] > 0 else 0
                low = int(np.searchsorted(c1[row], r - before, side="right"))
                bits[j] = (int(hi[j]) << 16) | low
                below[j] = before + (int(c1[row][low - 1]) if low > 0 else 0)
    return bits, below


def order_statistics(keys: torch.Tensor, ranks: Sequence[int], group=None) -> np.ndarray:
    """Exact values of the given 0-based ranks of the (globally sorted, non-negative) float32 or float64
    keys.  Two 16-bit radix-histogram passes on the device per 32 key bits; with `group` the histograms are
    summed over ranks (torch.distributed all-reduce) so every rank gets the global answer."""
    if keys.dtype == torch.float64:
        return _order_statistics_f64(keys, ranks, group)
    keys = _need_cuda("keys", keys, torch.float32, 1)
    bits, _ = _order_statistics_bits(keys, ranks, group)
    return bits.view(np.float32)


def _order_statistics_f64(keys: torch.Tensor, ranks: Sequence[int], group=None) -> np.ndarray:
    """float64 keys (calibrated confidences: isotonic / density-ratio outputs are float64 and differ from each other
    far below float32 resolution): the order statistic's upper 32 bits are located first, then its lower 32 bits
    among the keys that share them - the same radix kernels on each half, so the edges are exact in float64."""
    keys = _need_cuda("keys", keys, torch.float64, 1)
    b64 = keys.view(torch.int64)
    hi = (b64 >> 32).to(torch.int32)
    lo = (b64 & 0xFFFFFFFF).to(torch.int32)                      # wraps: the bit pattern is what the kernels read
    ranks = [int(r) for r in ranks]
    hbits, hbelow = _order_statistics_bits(hi.view(torch.float32), ranks, group)
    out = np.empty(len(ranks), np.uint64)
    for h in sorted(set(int(x) for x in hbits)):
        js = [j for j in range(len(ranks)) if int(hbits[j]) == h]
        sub = lo[hi == int(np.int32(np.uint32(h)))].contiguous()
        lbits, _ = _order_statistics_bits(sub.view(torch.float32), [ranks[j] - int(hbelow[j]) for j in js], group)
        for j, lb in zip(js, lbits):
            out[j] = (np.uint64(h) << np.uint64(32)) | np.uint64(lb)
    return out.view(np.float64)
