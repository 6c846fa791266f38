"""Host-side arithmetic on the tiny bin tables the kernels emit (a few dozen integers).

The O(N) work - binning, counting, summing, order statistics - happens on the GPU; what is left
here is O(n_bins) float64 arithmetic that turns an (n+1)-bin table into the reference's numbers:

  ECE          tools/metrics.py:90-130   (reference)
  MCE          tools/metrics.py:181-208
  AdaptiveECE  tools/metrics.py:212-236  (+ scikit-learn KBinsDiscretizer quantile edges)
  PIECE        tools/metrics.py:132-178
  accuracy / mean confidence   evaluators/vl_evaluator.py:70-84
"""
from __future__ import annotations

import numpy as np

FX_SCALE = float(1 << 40)


def uniform_thresholds(n_bins: int) -> np.ndarray:
    """Thresholds of the (n+1)-bin ECE table: np.linspace(0,1,n+1)[1:] as float64
    (tools/metrics.py:104).  Bin n is the overflow bin conf >= 1.0."""
    return np.linspace(0, 1, int(n_bins) + 1)[1:]


def _cols(table):
    t = np.asarray(table).view(np.uint64).reshape(-1, 3)
    cnt = t[:, 0].astype(np.float64)
    cor = t[:, 1].astype(np.float64)
    # exact integer -> float64 conversion of sums up to 2^64 loses < 2^-52 relative: irrelevant
    sm = t[:, 2].astype(np.float64) / FX_SCALE
    return cnt, cor, sm


def total_count(table) -> int:
    return int(np.asarray(table).view(np.uint64).reshape(-1, 3)[:, 0].sum())


def accuracy(table) -> float:
    cnt, cor, _ = _cols(table)
    return float(cor.sum() / cnt.sum())


def mean_confidence(table) -> float:
    cnt, _, sm = _cols(table)
    return float(sm.sum() / cnt.sum())


def reliability_bins(table) -> dict:
    """Per-bin accuracy / mean confidence / count of an (n+1)-bin uniform table, i.e. the data behind
    the reference's reliability diagram (tools/plot.py:8-71), overflow bin folded into the last bar."""
    cnt, cor, sm = _cols(table)
    n = len(cnt) - 1
    cnt, cor, sm = cnt.copy(), cor.copy(), sm.copy()
    cnt[n - 1] += cnt[n]; cor[n - 1] += cor[n]; sm[n - 1] += sm[n]
    cnt, cor, sm = cnt[:n], cor[:n], sm[:n]
    with np.errstate(invalid="ignore", divide="ignore"):
        acc = np.where(cnt > 0, cor / cnt, 0.0)
        conf = np.where(cnt > 0, sm / cnt, 0.0)
    edges = np.linspace(0, 1, n + 1)
    return {"edges": edges, "count": cnt.astype(np.int64), "accuracy": acc, "confidence": conf, "gap": conf - acc}


def ece_from_table(table) -> np.float64:
    """Reference ECE from the (n+1)-bin table built with uniform_thresholds(n).

    The reference takes the per-bin means over half-open bins [lo, hi) from np.digitize, so a
    confidence of exactly 1.0 is in NO bin mean, but takes the bin WEIGHTS from np.histogram
    whose last bin is closed, so that sample still adds to the last bin's weight
    (tools/metrics.py:105-127).  Hence: means from bins 0..n-1, overflow count folded into the
    weight of bin n-1 only."""
    cnt, cor, sm = _cols(table)
    n = len(cnt) - 1
    total = cnt.sum()
    if total == 0:
        return np.float64(np.nan)
    with np.errstate(invalid="ignore", divide="ignore"):
        acc = np.where(cnt[:n] > 0, cor[:n] / cnt[:n], 0.0)
        mc = np.where(cnt[:n] > 0, sm[:n] / cnt[:n], 0.0)
    w = cnt[:n].copy()
    w[n - 1] += cnt[n]
    return np.float64(np.sum(w / total * np.abs(mc - acc)))


def grouped_gaps(cnt, cor, sm) -> np.ndarray:
    total = cnt.sum()
    nz = cnt > 0
    return np.abs(cor[nz] / cnt[nz] - sm[nz] / cnt[nz]) * cnt[nz] / total


def mce_from_table(table) -> np.float64:
    """Reference MCE: inner edges only, so conf == 1.0 belongs to the last bin (overflow folded
    in completely), and the max of COUNT-WEIGHTED gaps (tools/metrics.py:198-206)."""
    cnt, cor, sm = _cols(table)
    n = len(cnt) - 1
    cnt, cor, sm = cnt.copy(), cor.copy(), sm.copy()
    cnt[n - 1] += cnt[n]; cor[n - 1] += cor[n]; sm[n - 1] += sm[n]
    g = grouped_gaps(cnt[:n], cor[:n], sm[:n])
    return np.float64(g.max()) if g.size else np.float64(np.nan)


def sum_of_gaps(table) -> np.float64:
    """AdaptiveECE / PIECE: sum over non-empty groups of |acc - mean conf| * count / N
    (tools/metrics.py:231-234, :164-168).  Works for 1-D and 2-D tables."""
    cnt, cor, sm = _cols(table)
    return np.float64(grouped_gaps(cnt, cor, sm).sum())


# --------------------------------------------------------------------------------------
# quantile bin edges = sklearn KBinsDiscretizer(strategy='quantile') on a float32 column
# --------------------------------------------------------------------------------------
def quantile_ranks(n: int, n_bins: int, method: str = "averaged_inverted_cdf"):
    """Which order statistics np.percentile(x, linspace(0,100,n_bins+1), method=...) needs.

    Returns (lo_rank, hi_rank, gamma) arrays of length n_bins+1; edge = lerp(x[lo], x[hi], gamma).
    Follows numpy's _quantile: virtual index = n*q - 1 (inverted-CDF family) or (n-1)*q
    (linear), evaluated in float64 exactly as numpy does."""
    q = np.true_divide(np.linspace(0, 100, int(n_bins) + 1), 100)
    if method in ("averaged_inverted_cdf", "inverted_cdf"):
        virt = n * q - 1
    elif method == "linear":
        virt = (n - 1) * q
    else:
        raise ValueError(f"unsupported quantile method {method!r}")
    prev = np.floor(virt)
    gamma = virt - prev
    if method == "averaged_inverted_cdf":
        gamma = np.where(gamma == 0, 0.5, 1.0)
    elif method == "inverted_cdf":
        gamma = np.where(gamma == 0, 0.0, 1.0)
    lo = np.clip(prev, 0, n - 1).astype(np.int64)
    hi = np.clip(prev + 1, 0, n - 1).astype(np.int64)
    # numpy clips the virtual index: below 0 everything collapses onto x[0]
    gamma = np.where(virt < 0, 0.0, gamma)
    hi = np.where(virt < 0, lo, hi)
    return lo, hi, gamma


def lerp_like_numpy(a: np.ndarray, b: np.ndarray, t: np.ndarray) -> np.ndarray:
    """numpy.lib._function_base_impl._lerp with numpy's own dtype behaviour: for a float32 column
    the difference b - a is taken in float32, the interpolation itself in float64."""
    a = np.asarray(a)
    b = np.asarray(b)
    t = np.asarray(t, dtype=np.float64)
    diff = np.subtract(b, a)
    out = np.asarray(np.add(a, diff * t), dtype=np.float64)
    return np.where(t >= 0.5, np.subtract(b, diff * (1 - t)), out)


def edges_from_order_stats(x_lo, x_hi, gamma) -> np.ndarray:
    """float64 bin edges with sklearn's 'drop edges closer than 1e-8' rule."""
    edges = np.asarray(lerp_like_numpy(x_lo, x_hi, gamma), dtype=np.float64)
    keep = np.ediff1d(edges, to_begin=np.inf) > 1e-8
    return edges[keep]


def macro_f1_from_counts(counts) -> float:
    """sklearn.metrics.f1_score(labels, preds, average="macro", labels=np.unique(labels)) (the call of
    evaluators/vl_evaluator.py:74-79) from a per-class {tp, fp, fn} table: F1 of every class that occurs among
    the true labels (tp + fn > 0), 0 where a class is never predicted correctly, averaged with equal weights."""
    counts = np.asarray(counts).astype(np.float64).reshape(-1, 3)
    tp, fp, fn = counts[:, 0], counts[:, 1], counts[:, 2]
    present = (tp + fn) > 0
    if not present.any():
        return 0.0
    tp, fp, fn = tp[present], fp[present], fn[present]
    pred_sum, true_sum = tp + fp, tp + fn
    with np.errstate(invalid="ignore", divide="ignore"):
        precision = np.where(pred_sum > 0, tp / pred_sum, 0.0)
        recall = tp / true_sum
        denom = precision + recall
        f1 = np.where(denom > 0, 2.0 * precision * recall / denom, 0.0)
    return float(np.mean(f1))
