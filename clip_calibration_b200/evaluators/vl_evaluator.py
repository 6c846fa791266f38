"""The metric half of the reference's evaluators/vl_evaluator.py (`VLClassification.evaluate`,
reference :59-116): accuracy, error, macro-F1, mean confidence, ECE, MCE, ACE (and PIECE when a
proximity vector is given), as percentages in the reference's result keys.

In scope: argmax / confidence gather / accuracy / mean confidence (:65-84), macro-F1 from per-class
{tp, fp, fn} counted on the device (:74-79) and the calibration metrics (:86-92).  The reliability plot
(:118-137, matplotlib) is tools/plot.py; the per-bin table it draws is returned under "bin_table".
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from .. import native
from .. import table_math as tm
from ..tools import metrics


def results_from_tables(table, class_counts) -> "OrderedDict[str, float]":
    """accuracy / error_rate / macro_f1 / confidence / ece / mce (reference keys and units) from the (n+1)-bin table
    and the per-class {tp, fp, fn} counts - everything that needs no second pass over the images."""
    results = OrderedDict()
    acc = 100.0 * tm.accuracy(table)
    results["accuracy"] = acc
    results["error_rate"] = 100.0 - acc
    results["macro_f1"] = 100.0 * tm.macro_f1_from_counts(class_counts)
    results["confidence"] = tm.mean_confidence(table)
    results["ece"] = 100.0 * float(tm.ece_from_table(table))
    results["mce"] = 100.0 * float(tm.mce_from_table(table))
    return results


def evaluate_pred_conf(preds, confs, labels, ece_bins: int = 10, piece_bins: int = 10, proximity=None,
                       group=None, n_classes=None) -> "OrderedDict[str, float]":
    """Metrics from per-image (pred, conf, label); numpy or CUDA tensors."""
    table = metrics.bin_stats(confs, preds, labels, ece_bins, group)
    results = results_from_tables(table, metrics.class_counts(preds, labels, n_classes, group))
    results["ace"] = 100.0 * float(metrics.AdaptiveECE(confs, preds, labels, ece_bins, group=group))
    if proximity is not None:
        results["piece"] = 100.0 * float(metrics.PIECE(confs, proximity, preds, labels, piece_bins, ece_bins,
                                                       group=group))
    results["bin_table"] = table
    return results


def evaluate(probs, labels, text_proximity=None, ece_bins: int = 10, piece_bins: int = 10):
    """Reference signature: a full probability matrix in, the result dict out."""
    p = probs if isinstance(probs, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(probs))
    if p.dtype == torch.float64:
        # calibrated probabilities (isotonic / density-ratio outputs) are float64 and may differ below float32
        # resolution: keep them - first argmax and its float64 value (torch: first maximal index), float64 binning
        p = p.to(device="cuda").contiguous()
        pred = torch.argmax(p, dim=1)
        conf = p.gather(1, pred[:, None])[:, 0].contiguous()
        pred = pred.to(torch.int32)
    else:
        p = p.to(device="cuda", dtype=torch.float32).contiguous()
        pred, conf = native.row_argmax(p)          # preds = argmax(probs), confs = probs[i, preds_i]
    return evaluate_pred_conf(pred, conf, labels, ece_bins, piece_bins, text_proximity, n_classes=p.shape[1])
