"""`HistogramBinning` and `IsotonicRegression` with the call surface `VLCalibration` uses from `netcal.binning`
(reference trainers/calibration/vl_calibrator.py:20-21; constructed at :125-131 under BinMeanShift and at :137-143
plain; called as `.fit(val_probs, val_labels)`, `.fit_transform(probs, labels)`, `.transform(probs)`).

netcal is a pip dependency of the reference (requirements.txt:6, unpinned) and is NOT vendored under the reference
tree, so these classes restate netcal 1.3's published algorithm; **parity with netcal itself is unpinned** (the package
is not installed here).  What is pinned: the binary isotonic fit inside is scikit-learn's, checked against scikit-learn
bit for knot; the CUDA path is checked against a numpy / scipy / scikit-learn statement of the scheme below (the
tests' CPU checker) and against the reference's own BinMeanShift run around that statement (tests/golden).

The scheme (netcal.AbstractCalibration, multi-class, `detection=False`):
  * fit: for every class j that occurs in y, a BINARY calibrator is fitted on the one-vs-all problem
    (X[:, j], y == j); classes without a sample get no calibrator (their column transforms to 0);
  * binary HistogramBinning(bins, equal_intervals=True): edges = np.linspace(0, 1, bins + 1); the bin map is the mean
    of the binary target per bin (scipy.stats.binned_statistic_dd, last bin closed); an empty bin maps to its centre;
  * binary IsotonicRegression: sklearn.isotonic.IsotonicRegression(increasing=True, out_of_bounds='clip');
  * transform: column j through calibrator j, then every row divided by its sum (`independent_probabilities=False`).
Arithmetic is float64: float32 probabilities are widened exactly, i.e. the result is what netcal returns for
`X.astype(np.float64)`.  Not provided: `detection=True`, `equal_intervals=False`, [N, 2] inputs (pass the positive
class's confidence as a 1-D array for a binary problem).  There is no CPU path.
"""
from __future__ import annotations

import numpy as np
import torch

from ... import native


def _probs_device(x) -> torch.Tensor:
    t = x.detach() if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float32 if t.dtype in (torch.float16, torch.bfloat16) else torch.float64)
    return t.cuda().contiguous()


def _labels_device(y, n: int) -> torch.Tensor:
    t = y.detach() if isinstance(y, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(y))
    t = t.cuda()
    if t.dim() == 2:                                 # one-hot targets: netcal takes their argmax
        t = t.argmax(dim=1)
    t = t.reshape(-1).to(torch.int64)
    if t.numel() != n:
        raise ValueError("one label per sample is required")
    return t.contiguous()


class _OneVsAll:
    """Shared plumbing: shapes, the binary (1-D) case, numpy in -> numpy out."""

    def __init__(self, detection: bool = False, independent_probabilities: bool = False):
        if detection:
            raise NotImplementedError("detection=True (box-conditioned calibration) is not provided")
        self.detection = detection
        self.independent_probabilities = independent_probabilities
        self.num_classes = None
        self._binary = False

    # --- to be provided: _fit_device(probs [N, C], labels [N]) and _apply_device(probs [N, C], normalise)
    def _shape(self, x: torch.Tensor, fitting: bool) -> torch.Tensor:
        if x.dim() == 2 and x.shape[1] == 1:
            x = x.reshape(-1)
        if x.dim() == 1:
            binary = True
        elif x.dim() == 2 and x.shape[1] >= 3:
            binary = False
        else:
            raise NotImplementedError("expected confidences [N] (binary) or probabilities [N, C] with C >= 3; for two "
                                      "classes pass the positive class's confidence as a 1-D array")
        if fitting:
            self._binary = binary
            self.num_classes = 2 if binary else x.shape[1]
        elif binary != self._binary or (not binary and x.shape[1] != self.num_classes):
            raise ValueError("transform input does not have the shape the calibrator was fitted on")
        return x.reshape(-1, 1) if binary else x

    def fit_device(self, probs: torch.Tensor, labels: torch.Tensor) -> "_OneVsAll":
        x = self._shape(probs, True)
        if self._binary:                             # the single "class 0" of the [N, 1] problem is the positive class
            labels = torch.where(labels == 1, 0, -1)
        self._fit_device(x, labels)
        return self

    def transform_device(self, probs: torch.Tensor) -> torch.Tensor:
        if self.num_classes is None:
            raise RuntimeError(f"{type(self).__name__}.transform called before fit")
        x = self._shape(probs, False)
        out = self._apply_device(x, normalise=not (self._binary or self.independent_probabilities))
        return out.reshape(-1) if self._binary else out

    def fit_transform_device(self, probs: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        return self.fit_device(probs, labels).transform_device(probs)

    def fit(self, X, y, random_state=None, tensorboard=None, log_dir=None):
        x = _probs_device(X)
        return self.fit_device(x, _labels_device(y, x.shape[0]))

    def transform(self, X):
        as_numpy = not isinstance(X, torch.Tensor)
        out = self.transform_device(_probs_device(X))
        return out.cpu().numpy() if as_numpy else out

    def fit_transform(self, X, y=None, **fit_params):
        return self.fit(X, y).transform(X)


class HistogramBinning(_OneVsAll):
    """netcal.binning.HistogramBinning(bins=10, equal_intervals=True): per-class accuracy histogram over
    equal-width confidence bins, one launch over the [N, C] matrix for all C one-vs-all problems."""

    def __init__(self, bins=10, equal_intervals=True, detection=False, independent_probabilities=False):
        super().__init__(detection, independent_probabilities)
        if not equal_intervals:
            raise NotImplementedError("equal_intervals=False (equal-frequency bins) is not provided")
        if isinstance(bins, (tuple, list)):
            if len(bins) != 1:
                raise NotImplementedError("one bin count is expected (confidence is the only feature)")
            bins = bins[0]
        self.bins = int(bins)
        self._bin_bounds = None                      # [np.linspace(0, 1, bins + 1)], like netcal's list per dimension
        self._bin_map = None                         # float64 [C, bins] (netcal keeps one [bins] map per sub-model)
        self._fitted_classes = None                  # bool [C]: classes that had a sample

    def _fit_device(self, x, labels):
        edges = np.linspace(0.0, 1.0, self.bins + 1)
        self._bin_bounds = [edges]
        self._edges_dev = torch.from_numpy(edges).cuda()
        count, hits = native.ova_hist_fit(x, labels, self._edges_dev)
        count, hits = count.cpu().numpy(), hits.cpu().numpy()
        centres = (edges[1:] + edges[:-1]) * 0.5
        with np.errstate(invalid="ignore", divide="ignore"):
            bin_map = np.where(count > 0, hits / count, centres[None, :])
        seen = hits.sum(axis=1) > 0
        if not self._binary:
            bin_map[~seen] = 0.0                     # no sample of the class: no sub-model, the column stays 0
        self._fitted_classes = seen
        self._bin_map = bin_map
        self._bin_map_dev = torch.from_numpy(np.ascontiguousarray(bin_map)).cuda()

    def _apply_device(self, x, normalise):
        return native.ova_apply(x, edges=self._edges_dev, bin_map=self._bin_map_dev, normalise=normalise)


class IsotonicRegression(_OneVsAll):
    """netcal.binning.IsotonicRegression: scikit-learn's isotonic fit per one-vs-all problem (GPU fit of
    csrc/isotonic.cu per class), all classes transformed by one launch."""

    _CLASS_CHUNK_BYTES = 1 << 30

    def __init__(self, detection=False, independent_probabilities=False):
        super().__init__(detection, independent_probabilities)
        self.knots = None                            # per class: (X_thresholds_, y_thresholds_) or None

    def _fit_device(self, x, labels):
        n, c = x.shape
        chunk = max(1, min(c, self._CLASS_CHUNK_BYTES // max(1, 9 * n)))
        kxs, kys, off, knots = [], [], [0], []
        for j0 in range(0, c, chunk):
            j1 = min(c, j0 + chunk)
            cols = x[:, j0:j1].t().to(torch.float64).contiguous()                  # [chunk, N]: one row per class
            ids = torch.arange(j0, j1, device=x.device, dtype=torch.int64)
            target = (labels.reshape(1, -1) == ids.reshape(-1, 1)).to(torch.uint8).contiguous()
            present = target.any(dim=1).cpu().numpy()
            for j in range(j1 - j0):
                if not present[j]:
                    knots.append(None)
                    off.append(off[-1])
                    continue
                kx, ky = native.isotonic_fit_binary(cols[j], target[j])
                kxs.append(kx)
                kys.append(ky)
                off.append(off[-1] + kx.numel())
                knots.append((kx.cpu().numpy(), ky.cpu().numpy()))
        self.knots = knots
        z = torch.zeros(1, dtype=torch.float64, device=x.device)
        self._kx = torch.cat(kxs) if kxs else z
        self._ky = torch.cat(kys) if kys else z
        self._off = torch.tensor(off, dtype=torch.int32, device=x.device)

    def _apply_device(self, x, normalise):
        return native.ova_apply(x, knots_x=self._kx, knots_y=self._ky, knot_off=self._off, normalise=normalise)
