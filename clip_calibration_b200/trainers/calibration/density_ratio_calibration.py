"""Drop-in for `DensityRatioCalibration` of the reference's trainers/calibration/density_ratio_calibration.py
(:28-117): same class name, `fit(probs, preds, true, proximity, bandwidth='normal_reference')`,
`predict(probs, proximities)` -> float64 probabilities [N, C], attributes `dens_true`, `dens_false`,
`false_true_ratio`.

The reference builds two statsmodels `KDEMultivariate` objects over (confidence, proximity) - one from the correctly
classified validation samples, one from the misclassified ones - and evaluates both with a Python loop over the test
points.  Here the densities are evaluated by one CUDA kernel (ccal_kde2_pdf: every (test point, validation point) pair
in parallel, float64 exponents, log-sum-exp range) and the probability rows are rewritten by a second one
(ccal_density_ratio_apply).  statsmodels is not needed; its `normal_reference` rule of thumb
(bw_j = 1.06 * std_j * nobs ** (-1/6) for two variables) is restated in `normal_reference_bandwidth`.
There is no CPU path: without an sm_100 GPU these calls raise.

The other classes of the reference file (`CustomizedDensityRatioCalibration`, the mirror_* helpers) have no caller in
the reference and are not provided.
"""
from __future__ import annotations

import numpy as np
import torch

from ... import native


def normal_reference_bandwidth(columns: np.ndarray) -> np.ndarray:
    """statsmodels `GenericKDE._normal_reference` for data [nobs, k_vars] (float64)."""
    nobs, k_vars = columns.shape
    return 1.06 * np.std(columns, axis=0) * nobs ** (-1.0 / (4 + k_vars))


class GaussianProductKDE:
    """The slice of `sm.nonparametric.KDEMultivariate(data=[dep, indep], var_type='cc', bw=...)` the reference uses:
    `.data` [nobs, 2] float64, `.bw` [2], `.nobs`, `.pdf(points [n, 2])` -> float64 [n] (numpy in, numpy out;
    `pdf_device` keeps everything on the GPU)."""

    def __init__(self, data, var_type="cc", bw="normal_reference"):
        if var_type != "cc":
            raise ValueError("only two continuous variables (var_type='cc') are supported")
        cols = np.asarray([np.asarray(v, dtype=np.float64) for v in data])
        if cols.ndim != 2 or cols.shape[0] != 2:
            raise ValueError("data must be [dep, indep]: two equally long 1-D arrays")
        self.data = np.ascontiguousarray(cols.T)
        self.nobs, self.k_vars = self.data.shape
        if self.nobs < 1:
            raise ValueError("the density needs at least one sample")
        if isinstance(bw, str):
            if bw != "normal_reference":
                raise ValueError("bandwidth rule %r is not supported (only 'normal_reference' or explicit values)" % bw)
            self.bw = normal_reference_bandwidth(self.data)
        else:
            self.bw = np.asarray(bw, dtype=np.float64).reshape(2)
        if not np.all(self.bw > 0):
            raise ValueError(f"degenerate bandwidth {self.bw}: the samples have no spread in one variable")
        dev = torch.from_numpy(self.data).cuda()
        self._x, self._y = dev[:, 0].contiguous(), dev[:, 1].contiguous()

    def pdf_device(self, qx: torch.Tensor, qy: torch.Tensor) -> torch.Tensor:
        return native.kde2_pdf(self._x, self._y, qx, qy, float(self.bw[0]), float(self.bw[1]))

    def pdf(self, data_predict=None):
        pts = self.data if data_predict is None else np.asarray(data_predict, dtype=np.float64).reshape(-1, 2)
        q = torch.from_numpy(np.ascontiguousarray(pts)).cuda()
        return self.pdf_device(q[:, 0].contiguous(), q[:, 1].contiguous()).cpu().numpy()


def _as_device_probs(probs) -> torch.Tensor:
    if isinstance(probs, torch.Tensor):
        t = probs.detach()
    else:
        t = torch.from_numpy(np.ascontiguousarray(probs))
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float32 if t.dtype in (torch.float16, torch.bfloat16) else torch.float64)
    return t.cuda().contiguous()


class DensityRatioCalibration():

    def __init__(self):
        pass

    def fit(self, probs, preds, true, proximity, bandwidth="normal_reference"):
        """Densities of (confidence, proximity) for the correctly / wrongly classified samples, reference :35-72."""
        p = _as_device_probs(probs)
        lo, hi = float(p.min()), float(p.max())
        assert lo >= 0 and hi <= 1, "All elements in 'probs' should be in the range [0, 1]."
        confs = self._row_max(p).cpu().numpy()        # keeps the dtype of probs, like np.max(probs, axis=-1)
        correct = np.asarray(preds) == np.asarray(true)
        proximity = np.asarray(proximity)
        if correct.all() or not correct.any():
            raise ValueError("density-ratio calibration needs both correctly and wrongly classified validation samples")
        self.dens_true = GaussianProductKDE([confs[correct], proximity[correct]], var_type="cc", bw=bandwidth)
        self.dens_false = GaussianProductKDE([confs[~correct], proximity[~correct]], var_type="cc", bw=bandwidth)
        self.false_true_ratio = (~correct).sum() / correct.sum()

    @staticmethod
    def _row_max(p: torch.Tensor) -> torch.Tensor:
        return native.row_argmax(p)[1] if p.dtype == torch.float32 else p.max(dim=1).values

    def predict_device(self, probs: torch.Tensor, proximities: torch.Tensor):
        """CUDA tensors in and out: (probs_out float64 [N, C], conf_calibrated float64 [N], pred int32 [N])."""
        qx = self._row_max(probs).to(torch.float64)
        qy = proximities.to(device=probs.device, dtype=torch.float64).contiguous()
        t = self.dens_true.pdf_device(qx, qy)
        f = self.dens_false.pdf_device(qx, qy)
        return native.density_ratio_apply(probs, t, f, float(self.false_true_ratio))

    def predict(self, probs, proximities):
        """Bayes posterior of `correct` given (confidence, proximity) written into the probability rows, reference
        :78-117.  numpy in -> numpy float64 out; CUDA tensors in -> CUDA float64 tensor out."""
        as_numpy = not isinstance(probs, torch.Tensor)
        p = _as_device_probs(probs)
        lo, hi = float(p.min()), float(p.max())
        assert lo >= 0 and hi <= 1, "All elements in 'probs' should be in the range [0, 1]."
        prox = proximities if isinstance(proximities, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(proximities))
        out, _, _ = self.predict_device(p, prox.cuda())
        return out.cpu().numpy() if as_numpy else out

    def calibrated_confidence(self, conf, proximities):
        """Additive: only conf_calibrated [N] from per-sample confidences (no [N, C] matrix)."""
        as_numpy = not isinstance(conf, torch.Tensor)
        qx = (torch.from_numpy(np.ascontiguousarray(conf)) if as_numpy else conf).cuda().to(torch.float64).contiguous()
        qy = (torch.from_numpy(np.ascontiguousarray(proximities)) if not isinstance(proximities, torch.Tensor)
              else proximities).cuda().to(torch.float64).contiguous()
        t = self.dens_true.pdf_device(qx, qy)
        f = self.dens_false.pdf_device(qx, qy)
        # a one-column "probability matrix": the row kernel then only forms t / max(t + f * ratio, 1e-10)
        _, cal, _ = native.density_ratio_apply(qx.reshape(-1, 1), t, f, float(self.false_true_ratio))
        return cal.cpu().numpy() if as_numpy else cal
