"""Drop-in for the reference's tools/metrics.py (ECE / MCE / AdaptiveECE / PIECE), same names,
argument order, defaults and float64 *fraction* return values (the caller multiplies by 100,
evaluators/vl_evaluator.py:86-92).

The O(N) part - binning, per-bin counting and summing, the order statistics that define the
quantile bins - runs in CUDA (ccal_bin_stats / ccal_radix_hist); the host only combines an
(n_bins+1)-entry integer table.  numpy inputs are copied to the GPU, CUDA tensors are used in
place.  There is no CPU fallback: without an sm_100 GPU these functions raise.

Additive entry points (not in the reference): bin_stats, ece_from_table, mce_from_table,
calibration_summary, class_counts / macro_f1 (the evaluator's sklearn call), and the `group=` argument that all-reduces tables across ranks.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import native
from .. import table_math as tm
from ..table_math import ece_from_table, mce_from_table, uniform_thresholds  # noqa: F401  (re-export)


def _dev(x, dtype=None) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        t = x if x.is_cuda else x.cuda()
    else:
        t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous().reshape(-1)


def _conf_dev(conf) -> torch.Tensor:
    t = _dev(conf)
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float32 if t.dtype in (torch.float16, torch.bfloat16) else torch.float64)
    return t


def _pred_dev(pred) -> torch.Tensor:
    t = _dev(pred)
    return t if t.dtype in (torch.int32, torch.int64) else t.to(torch.int64)


def _allreduce(table: torch.Tensor, group) -> torch.Tensor:
    if group is not None:
        torch.distributed.all_reduce(table, group=group)
    return table


def bin_stats(conf, pred, gt, n_bins: int = 10, group=None) -> np.ndarray:
    """(n_bins+1, 3) uint64 table {count, n_correct, sum round(conf*2^40)} over the bins of
    np.digitize(conf, np.linspace(0,1,n_bins+1)) - 1 (last row = conf >= 1.0)."""
    table = native.bin_stats(_conf_dev(conf), _pred_dev(pred), _dev(gt, torch.int64), uniform_thresholds(n_bins))
    return native.table_to_numpy(_allreduce(table, group))


def ECE(conf, pred, gt, conf_bin_num=10, group=None):
    """Expected Calibration Error, reference tools/metrics.py:90-130."""
    return ece_from_table(bin_stats(conf, pred, gt, conf_bin_num, group))


def MCE(conf, pred, gt, conf_bin_num=10, group=None):
    """Maximal (count-weighted) Calibration Error, reference tools/metrics.py:181-208."""
    return mce_from_table(bin_stats(conf, pred, gt, conf_bin_num, group))


def quantile_thresholds(keys: torch.Tensor, n_bins: int, quantile_method: str = "averaged_inverted_cdf",
                        group=None) -> np.ndarray:
    """Inner bin edges of KBinsDiscretizer(n_bins, strategy='quantile') fitted on `keys`
    (float32 or float64, non-negative), from exact global order statistics computed on the device.
    Matches scikit-learn 1.9 for N <= 200,000; above that sklearn estimates the edges from an
    unseeded 200k subsample while this uses all N values."""
    n_local = keys.numel()
    n = n_local
    if group is not None:
        cnt = torch.tensor([n_local], dtype=torch.int64, device=keys.device)
        torch.distributed.all_reduce(cnt, group=group)
        n = int(cnt.item())
    lo, hi, gamma = tm.quantile_ranks(n, n_bins, quantile_method)
    ranks = np.unique(np.concatenate([lo, hi, [0, n - 1]]))
    vals = native.order_statistics(keys, ranks.tolist(), group=group)
    look = dict(zip(ranks.tolist(), vals))
    if look[0] == look[n - 1]:
        return np.zeros(0, np.float64)          # constant column: sklearn collapses to one bin
    dt = np.float64 if keys.dtype == torch.float64 else np.float32
    x_lo = np.array([look[int(r)] for r in lo], dt)
    x_hi = np.array([look[int(r)] for r in hi], dt)
    edges = tm.edges_from_order_stats(x_lo, x_hi, gamma)
    return edges[1:-1]


def uniform_thresholds_of(keys: torch.Tensor, n_bins: int, group=None) -> np.ndarray:
    """Inner bin edges of KBinsDiscretizer(n_bins, strategy='uniform') fitted on the float32 column `keys`:
    np.linspace(min, max, n_bins + 1)[1:-1] evaluated in float32 as scikit-learn 1.9 does (bin = number of edges <= x);
    empty for a constant column (one bin).  min / max come from the device (global over `group`)."""
    n = keys.numel()
    if group is not None:
        cnt = torch.tensor([n], dtype=torch.int64, device=keys.device)
        torch.distributed.all_reduce(cnt, group=group)
        n = int(cnt.item())
    if n == 0:
        return np.zeros(0, np.float64)
    lo, hi = native.order_statistics(keys, [0, n - 1], group=group)
    if lo == hi:
        return np.zeros(0, np.float64)
    return np.linspace(np.float32(lo), np.float32(hi), int(n_bins) + 1)[1:-1].astype(np.float64)


def AdaptiveECE(conf, pred, gt, conf_bin_num=10, quantile_method="averaged_inverted_cdf", group=None):
    """Equal-mass-bin ECE, reference tools/metrics.py:212-236."""
    c = _conf_dev(conf)
    # float64 confidences (isotonic / density-ratio outputs) keep their own resolution: edges from float64 order
    # statistics, binning of the float64 values - what KBinsDiscretizer does with a float64 column
    thr = quantile_thresholds(c, conf_bin_num, quantile_method, group)
    table = native.bin_stats(c, _pred_dev(pred), _dev(gt, torch.int64), thr)
    return tm.sum_of_gaps(native.table_to_numpy(_allreduce(table, group)))


def PIECE(conf, knndist, pred, gt, dist_bin_num=10, conf_bin_num=10, knn_strategy="quantile",
          quantile_method="averaged_inverted_cdf", group=None):
    """Proximity-informed ECE, reference tools/metrics.py:132-178: groups = (bin of knndist) x (uniform
    inner-edge bin of conf).  knn_strategy (reference :152, handed to KBinsDiscretizer): "quantile" (the default and
    the reference's only caller) = exact global order statistics; "uniform" = np.linspace(min, max, bins + 1) in
    float32 like scikit-learn on a float32 column (min / max are order statistics 0 and N-1, global over `group`).
    "kmeans" (scikit-learn's 1-D Lloyd iterations) is not provided."""
    if knn_strategy not in ("quantile", "uniform"):
        raise ValueError(f"knn_strategy={knn_strategy!r}: 'quantile' and 'uniform' are supported; 'kmeans' "
                         "(scikit-learn's KMeans on the proximity column) is not")
    key2 = _dev(knndist, torch.float32)
    if knn_strategy == "quantile":
        thr2 = quantile_thresholds(key2, dist_bin_num, quantile_method, group)
    else:
        thr2 = uniform_thresholds_of(key2, dist_bin_num, group)
    thr = np.linspace(0, 1, int(conf_bin_num) + 1)[1:-1]
    if len(thr2) == 0:
        table = native.bin_stats(_conf_dev(conf), _pred_dev(pred), _dev(gt, torch.int64), thr)
    else:
        table = native.bin_stats(_conf_dev(conf), _pred_dev(pred), _dev(gt, torch.int64), thr, key2, thr2)
    return tm.sum_of_gaps(native.table_to_numpy(_allreduce(table, group)))


def class_counts(pred, gt, n_classes=None, group=None) -> np.ndarray:
    """[n_classes, 3] int64 table {tp, fp, fn} per class, counted on the device (ccal_class_counts)."""
    p, g = _pred_dev(pred), _dev(gt, torch.int64)
    if n_classes is None:
        n_classes = int(max(int(p.max()), int(g.max())) + 1) if p.numel() else 1
        if group is not None:
            c = torch.tensor([n_classes], dtype=torch.int64, device=p.device)
            torch.distributed.all_reduce(c, op=torch.distributed.ReduceOp.MAX, group=group)
            n_classes = int(c.item())
    counts = native.class_counts(p, g, n_classes)
    return _allreduce(counts, group).cpu().numpy()


def macro_f1(pred, gt, n_classes=None, group=None) -> float:
    """`f1_score(labels, preds, average="macro", labels=np.unique(labels))` as a fraction
    (evaluators/vl_evaluator.py:74-79 multiplies by 100)."""
    return tm.macro_f1_from_counts(class_counts(pred, gt, n_classes, group))


def calibration_summary(table) -> dict:
    """accuracy, mean confidence, ECE and MCE (fractions) from one (n+1)-bin table."""
    return {"n": tm.total_count(table), "accuracy": tm.accuracy(table), "confidence": tm.mean_confidence(table),
            "ece": float(ece_from_table(table)), "mce": float(mce_from_table(table))}


def compute_acc_bin(conf_thresh_lower, conf_thresh_upper, conf, pred, true):
    """Accuracy, mean confidence and size of the bin (lower, upper] - the reference's unused
    helper (tools/metrics.py:33-55), kept for surface completeness; host-side."""
    conf = np.asarray(conf)
    sel = (conf > conf_thresh_lower) & (conf <= conf_thresh_upper)
    n = int(sel.sum())
    if n < 1:
        return 0, 0, 0
    correct = int((np.asarray(pred)[sel] == np.asarray(true)[sel]).sum())
    return float(correct) / n, float(np.sum(conf[sel], dtype=np.float64) / n), n
