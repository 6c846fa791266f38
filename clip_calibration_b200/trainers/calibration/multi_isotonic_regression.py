"""Drop-in for `MultiIsotonicRegression` of the reference's trainers/calibration/multi_isotonic_regression.py
(:6-35): same class name, `fit_transform(logit, label)` and `transform(logit)` returning float64 [N, C],
`.calibrator` carrying scikit-learn's fitted attributes (`X_thresholds_`, `y_thresholds_`, `X_min_`, `X_max_`).

The reference flattens the [N, C] matrix and runs scikit-learn's IsotonicRegression (sort + pool-adjacent-violators
in Cython, then scipy interp1d) on N*C points.  Here the same fit runs on the GPU: float64 exp-normalisation,
radix sort, pooling in parallel rounds on exact integer (ones, count) blocks, interpolation (csrc/isotonic.cu).
There is no CPU path: without an sm_100 GPU these calls raise.  Targets must be class labels (1-D integers) or a
0/1 matrix - which is all the reference feeds it.
"""
from __future__ import annotations

import numpy as np
import torch

from ... import native


class _FittedIsotonic:
    """What the reference reads from `self.calibrator` (sklearn.isotonic.IsotonicRegression) after fitting."""
    out_of_bounds = "clip"
    increasing = True

    def __init__(self, knots_x: torch.Tensor, knots_y: torch.Tensor):
        self.knots_x, self.knots_y = knots_x, knots_y
        self.X_thresholds_ = knots_x.cpu().numpy()
        self.y_thresholds_ = knots_y.cpu().numpy()
        self.X_min_, self.X_max_ = self.X_thresholds_[0], self.X_thresholds_[-1]

    def predict_device(self, t: torch.Tensor, residual_scale: float = 0.0) -> torch.Tensor:
        return native.isotonic_transform(self.knots_x, self.knots_y, t, residual_scale)

    def predict(self, T):
        t = torch.from_numpy(np.ascontiguousarray(T, dtype=np.float64)).cuda()
        return self.predict_device(t).cpu().numpy()

    transform = predict


def _device_matrix(x) -> torch.Tensor:
    t = x.detach() if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
    if t.dim() != 2:
        raise ValueError(f"expected an [N, C] matrix, got shape {tuple(t.shape)}")
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float32 if t.dtype in (torch.float16, torch.bfloat16) else torch.float64)
    return t.cuda().contiguous()


class MultiIsotonicRegression():
    """multi-class isotonic regression (Mix-n-Match), reference multi_isotonic_regression.py:6-35."""

    def __init__(self) -> None:
        self.__name__ = 'MultiIsotonicRegression'
        self.calibrator = None

    def fit_transform_device(self, logit: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
        n, c = logit.shape
        if label.dim() == 1:
            # label_binarize(label, classes=arange(C)) / the two-class branch (:17-24): one-hot on the device
            p, onehot = native.exp_normalise_rows(logit, label.to(device=logit.device, dtype=torch.int64))
        else:
            if tuple(label.shape) != (n, c):
                raise ValueError("a label matrix must have the shape of logit")
            p, _ = native.exp_normalise_rows(logit)
            lab = label.to(device=logit.device)
            if not bool(((lab == 0) | (lab == 1)).all()):
                raise ValueError("label matrix must be one-hot (0/1): only binary targets are supported")
            onehot = lab.to(torch.uint8).contiguous()
        kx, ky = native.isotonic_fit_binary(p, onehot)
        self.calibrator = _FittedIsotonic(kx, ky)
        # y_ = calibrator.fit_transform(p.flatten(), label.flatten());  p = y_.reshape(...) + 1e-9 * p   (:27-28)
        return self.calibrator.predict_device(p, 1e-9)

    def transform_device(self, logit: torch.Tensor) -> torch.Tensor:
        if self.calibrator is None:
            raise RuntimeError("MultiIsotonicRegression.transform called before fit_transform")
        p, _ = native.exp_normalise_rows(logit)
        return self.calibrator.predict_device(p, 1e-9)

    def fit_transform(self, logit, label):
        as_numpy = not isinstance(logit, torch.Tensor)
        lab = label if isinstance(label, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(label))
        out = self.fit_transform_device(_device_matrix(logit), lab)
        return out.cpu().numpy() if as_numpy else out

    def transform(self, logit):
        as_numpy = not isinstance(logit, torch.Tensor)
        out = self.transform_device(_device_matrix(logit))
        return out.cpu().numpy() if as_numpy else out
