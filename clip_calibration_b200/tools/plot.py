"""Drop-in for `plot_reliability_diagram` of the reference's tools/plot.py (:8-71), the one plot function the
reference calls (evaluators/vl_evaluator.py:137): same name, arguments and figure, but the per-bin accuracy, mean
confidence and weights come from the device bin table (ccal_bin_stats) instead of a numpy pass per bin.

`reliability_diagram_data` is the additive, matplotlib-free half: exactly the arrays the reference computes at
:11-36 (bins by np.digitize -> a confidence of 1.0 is in no bin mean but counts in the last bin's weight).
matplotlib is imported only when a figure is requested; without it `plot_reliability_diagram` raises ImportError.
The exploratory plots of the reference file that nothing calls (`plot_proximity_conf`, `plot_proximity_acc_ece`,
`reliability_diagram`, `compute_ece`) are not provided.
"""
from __future__ import annotations

import numpy as np

from .. import table_math as tm
from . import metrics


def reliability_diagram_data(preds, confs, labels, n_bins=15, group=None, table=None) -> dict:
    """bin_acc [n], bin_confidences [n], weights [n], ece - the numbers behind the figure (reference :11-36).
    Pass `table` (an (n_bins+1)-row bin table, e.g. CalibratedScorer.reduced_table()) to skip the device pass."""
    if table is None:
        table = metrics.bin_stats(confs, preds, labels, n_bins, group)
    cnt, cor, sm = tm._cols(table)
    n = len(cnt) - 1
    with np.errstate(invalid="ignore", divide="ignore"):
        bin_acc = np.where(cnt[:n] > 0, cor[:n] / cnt[:n], 0.0)
        bin_conf = np.where(cnt[:n] > 0, sm[:n] / cnt[:n], 0.0)
    w = cnt[:n].astype(np.float64).copy()
    w[n - 1] += cnt[n]                                  # np.histogram's last bin is closed: conf == 1.0 counts here
    total = cnt.sum()
    weights = w / total if total > 0 else w
    return {"bin_acc": bin_acc, "bin_confidences": bin_conf, "weights": weights,
            "ece": float(np.sum(weights * np.abs(bin_conf - bin_acc))), "table": table}


def plot_reliability_diagram(preds, confs, labels, n_bins=15, title=None, save_dir=None, group=None, table=None):
    d = reliability_diagram_data(preds, confs, labels, n_bins, group, table)
    import matplotlib
    matplotlib.use("Agg", force=False)
    import matplotlib.pyplot as plt
    bin_acc, ece = d["bin_acc"], d["ece"]
    n_bins = len(bin_acc)
    delta = 1.0 / n_bins
    x = np.arange(0, 1, delta)[:n_bins]
    mid = np.linspace(delta / 2, 1 - delta / 2, n_bins)
    error = np.abs(np.subtract(mid, bin_acc))
    plt.rcParams["font.family"] = "serif"
    plt.figure(figsize=(6, 6))
    plt.xlim(0, 1)
    plt.ylim(0, 1)
    plt.grid(color="tab:grey", linestyle=(0, (1, 5)), linewidth=1, zorder=0)
    plt.bar(x, bin_acc, color="b", width=delta, align="edge", edgecolor="k", label="Outputs", zorder=5)
    plt.bar(x, error, bottom=np.minimum(bin_acc, mid), color="mistyrose", alpha=0.5, width=delta, align="edge",
            edgecolor="r", hatch="/", label="Gap", zorder=10)
    plt.plot([0.0, 1.0], [0.0, 1.0], linestyle="--", color="tab:grey", zorder=15)
    plt.ylabel("Accuracy", fontsize=13)
    plt.xlabel("Confidence", fontsize=13)
    plt.legend(loc="upper left", framealpha=1.0, fontsize="medium")
    plt.text(0.025, 0.85, f"ECE: {ece*100:.2f}%", transform=plt.gca().transAxes,
             bbox=dict(boxstyle="round, pad=0.5", facecolor="wheat", edgecolor="orange"))
    if title is not None:
        plt.title(title, fontsize=16)
    plt.tight_layout()
    if save_dir is not None:
        plt.savefig(save_dir)
    return plt
