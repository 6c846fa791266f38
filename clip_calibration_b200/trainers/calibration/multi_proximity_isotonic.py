"""Drop-in for `BinMeanShift` of the reference's trainers/calibration/multi_proximity_isotonic.py (:130-247), the
proximity-binned wrapper `VLCalibration` builds for the bin-based calibrators (vl_calibrator.py:121-135): samples are
grouped by proximity quantile (or uniform) bins and every group gets its own calibrator.

Provided: `method_name='multi_isotonic_regression'` with `MultiIsotonicRegression` (the scikit-learn based calibrator;
GPU fit, see multi_isotonic_regression.py), 'histogram_binning' / 'isotonic_regression' with the one-vs-all calibrators
of netcal_binning.py (netcal's published scheme; parity with netcal itself unpinned - it is not installed), `bin_strategy`
'quantile' (np.percentile edges from exact device order statistics) and 'uniform'.  Not provided: 'kmeans' bins, and
`MultiProximityIsotonicRegression`, which nothing in the reference calls.
"""
from __future__ import annotations

import numpy as np
import torch

from ... import native
from ... import table_math as tm
from .multi_isotonic_regression import MultiIsotonicRegression, _device_matrix


class BinMeanShift():

    def __init__(self, method_name, method, bin_strategy='quantile', normalize_conf=False, proximity_bin=10, **kwargs) -> None:
        if method_name not in ('multi_isotonic_regression', 'histogram_binning', 'isotonic_regression'):
            raise NotImplementedError(f"BinMeanShift method {method_name!r}: 'multi_isotonic_regression', "
                                      "'histogram_binning' and 'isotonic_regression' are provided")
        if bin_strategy not in ('quantile', 'uniform'):
            raise NotImplementedError(f"bin_strategy {bin_strategy!r}: only 'quantile' and 'uniform' are provided")
        self.method_name = method_name
        self.proximity_bin = proximity_bin
        self.bin_strategy = bin_strategy
        self.normalize_conf = normalize_conf
        self.calibrators = [method(**kwargs) for i in range(proximity_bin)]
        self.bin_edges = None

    # ------------------------------------------------------------------ bin edges (reference :158-184)
    @staticmethod
    def _prox_device(proximity) -> torch.Tensor:
        t = proximity.detach() if isinstance(proximity, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(proximity))
        return t.to(device="cuda", dtype=torch.float32).contiguous().reshape(-1)

    def get_bin_edges_by_quantile(self, proximity):
        """np.percentile(proximity, linspace(0, 100, bins + 1)) from exact order statistics (float32 proximities, as
        the reference pipeline produces them; the 'linear' method of np.percentile)."""
        keys = self._prox_device(proximity)
        n = keys.numel()
        lo, hi, gamma = tm.quantile_ranks(n, self.proximity_bin, "linear")
        ranks = np.unique(np.concatenate([lo, hi]))
        look = dict(zip(ranks.tolist(), native.order_statistics(keys, ranks.tolist())))
        x_lo = np.array([look[int(r)] for r in lo], np.float32)
        x_hi = np.array([look[int(r)] for r in hi], np.float32)
        return np.asarray(tm.lerp_like_numpy(x_lo, x_hi, gamma))

    def get_bin_edges_by_uniform(self, proximity):
        keys = self._prox_device(proximity)
        ext = native.order_statistics(keys, [0, keys.numel() - 1])
        return np.linspace(np.float32(ext[0]), np.float32(ext[1]), self.proximity_bin + 1)

    def _bin_numbers(self, proximity) -> torch.Tensor:
        """np.searchsorted(bin_edges[1:-1], proximity, side='right') on the device."""
        inner = torch.from_numpy(np.asarray(self.bin_edges[1:-1], dtype=np.float64)).cuda()
        return torch.searchsorted(inner, self._prox_device(proximity).to(torch.float64), right=True)

    # ------------------------------------------------------------------ reference :198-247
    def fit_transform(self, logit, proximity, label):
        if self.bin_strategy == 'quantile':
            self.bin_edges = self.get_bin_edges_by_quantile(proximity)
        else:
            self.bin_edges = self.get_bin_edges_by_uniform(proximity)
        return self._apply(logit, proximity, label)

    def transform(self, logit, proximity):
        if self.bin_edges is None:
            raise RuntimeError("BinMeanShift.transform called before fit_transform")
        return self._apply(logit, proximity, None)

    def _apply(self, logit, proximity, label):
        as_numpy = not isinstance(logit, torch.Tensor)
        x = _device_matrix(logit)
        if self.method_name in ('histogram_binning', 'isotonic_regression'):
            # logit = np.exp(logit) / np.sum(np.exp(logit), 1)[:, None]   (reference :219-220, :241-242)
            x, _ = native.exp_normalise_rows(x)
        bin_no = self._bin_numbers(proximity)
        if bin_no.numel() != x.shape[0]:
            raise ValueError("one proximity value per row is required")
        lab = None
        if label is not None:
            lab = (label if isinstance(label, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(label))).cuda()
        out = torch.empty(x.shape, dtype=torch.float64, device=x.device)
        for b in range(self.proximity_bin):
            idx = torch.nonzero(bin_no == b).reshape(-1)
            if idx.numel() == 0:
                if lab is not None:
                    raise ValueError(f"proximity bin {b} is empty: fewer distinct proximities than bins")
                continue
            rows = x.index_select(0, idx)
            if lab is not None:
                res = self.calibrators[b].fit_transform_device(rows, lab.index_select(0, idx))
            else:
                res = self.calibrators[b].transform_device(rows)
            out.index_copy_(0, idx, res)
        if self.normalize_conf and lab is not None:           # the reference normalises in fit_transform only (:225-226)
            out = out / out.sum(dim=1, keepdim=True)
        return out.cpu().numpy() if as_numpy else out
