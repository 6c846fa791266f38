"""CPU oracle for the scoring + calibration + metrics hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it.  Nothing under
clip_calibration_b200/ imports it and the product path never falls back to it.

It restates, in numpy/scipy/torch-CPU, the algorithm the reference runs for this path.
Every function cites the reference file:line it follows (paths relative to the reference
repo root).  Parity of this restatement against the real reference functions is pinned
by oracle/make_golden.py (run in the build container where /root/reference exists) and
re-checked by tests/test_oracle_golden.py against the committed fixtures.

Pinned: dac_fit, dac_predict, ece, mce, adaptive_ece, piece, knn_dists (against the
imported reference functions) and softmax / argmax / confidence (against scipy+numpy as the
reference calls them).  Parity UNPINNED: the TempScaling *trajectory* (dassl optimiser
semantics are not in the reference tree); ts_loss_and_grad is pinned to torch autograd only.
density_ratio_fit / density_ratio_predict: the reference's own arithmetic is pinned by running the real
DensityRatioCalibration class over KDEMultivariateCC; the KDE itself restates statsmodels (absent here,
unpinned in the reference's requirements.txt) -> that part is parity UNPINNED, cross-checked against sklearn.
NetcalHistogramBinningCC / NetcalIsotonicRegressionCC: netcal (requirements.txt:6, unpinned, absent here) restated from
netcal 1.3's published one-vs-all scheme -> parity with netcal UNPINNED; the reference's own BinMeanShift is run
unmodified over these classes by make_golden.py, which pins the reference-side arithmetic around them.
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------------------
# a-1  logit contraction
# --------------------------------------------------------------------------------------
def logits_fp32(img: np.ndarray, txt: np.ndarray, logit_scale: float, threads: int | None = None) -> np.ndarray:
    """`logits = logit_scale * image_features @ text_features.t()` in fp32 on CPU.

    Python precedence: the scale multiplies the image features first, then the GEMM
    (trainers/classification/zsclip.py:97-102, trainers/calibration/tempscaling.py:53-56).
    """
    import torch
    if threads is not None:
        torch.set_num_threads(int(threads))
    a = torch.from_numpy(np.ascontiguousarray(img, dtype=np.float32))
    b = torch.from_numpy(np.ascontiguousarray(txt, dtype=np.float32))
    s = torch.tensor(logit_scale, dtype=torch.float32)
    return (s * a @ b.t()).numpy()


# --------------------------------------------------------------------------------------
# a-6  DAC fit
# --------------------------------------------------------------------------------------
def _knn_sorted(ref_rows: np.ndarray, query: np.ndarray, k: int):
    # np.linalg.norm(rows - q, axis=1); np.sort(...)[:k]; np.argsort(...)[:k]
    # (trainers/calibration/distanse_aware_calibration.py:28-30 and :34-36)
    d = np.linalg.norm(ref_rows - query, axis=1)
    order = np.argsort(d, kind="stable")[:k]
    return d[order], order


def dac_fit(base_zs, cur_zs, base_tuned, cur_tuned, k: int):
    """Per-class logit multiplier (trainers/calibration/distanse_aware_calibration.py:13-46).

    Returns (class_confidence[C] float64, knn_idx_zs[C,kk], knn_idx_tuned[C,kk],
    knn_dist_zs[C,kk], knn_dist_tuned[C,kk]) with kk = min(k, B).  Quirks kept: the mean
    divides by k even if fewer than k base rows exist (:31, :37); the "< 0.05" base-class
    test looks at the TUNED nearest distance (:39, variable reused from :35); arithmetic
    stays in the input dtype.
    """
    C = cur_zs.shape[0]
    B = base_zs.shape[0]
    kk = min(k, B)
    cc = []
    idx_zs = np.zeros((C, kk), np.int64)
    idx_tu = np.zeros((C, kk), np.int64)
    d_zs = np.zeros((C, kk), base_zs.dtype)
    d_tu = np.zeros((C, kk), base_tuned.dtype)
    for i in range(C):
        dz, iz = _knn_sorted(base_zs, cur_zs[i], k)
        zs_score = np.exp(-np.sum(dz) / k)
        dt, it = _knn_sorted(base_tuned, cur_tuned[i], k)
        fs_score = np.exp(-np.sum(dt) / k)
        cc.append(1.0 if dt[0] < 0.05 else fs_score / zs_score)
        idx_zs[i], idx_tu[i], d_zs[i], d_tu[i] = iz, it, dz, dt
    return np.array(cc), idx_zs, idx_tu, d_zs, d_tu


# --------------------------------------------------------------------------------------
# a-7  DAC predict on materialised logits
# --------------------------------------------------------------------------------------
def dac_predict(logits: np.ndarray, class_confidence: np.ndarray) -> np.ndarray:
    """fp32: `pred = logits.max(1)[1]; logits *= cc[pred][:, None]`
    (trainers/calibration/distanse_aware_calibration.py:49-58; CPU statement :62-74).
    Tie-break: first maximum (numpy argmax)."""
    x = np.asarray(logits).astype(np.float32)
    cc = np.asarray(class_confidence).astype(np.float32)
    pred = np.argmax(x, axis=1)
    return x * cc[pred][:, None]


# --------------------------------------------------------------------------------------
# a-8 / a-9  softmax, argmax, confidence gather
# --------------------------------------------------------------------------------------
def softmax_lastaxis(x: np.ndarray) -> np.ndarray:
    """scipy.special.softmax(x, axis=-1) as called at trainers/calibration/vl_calibrator.py:91
    (shift by the row max, exp, divide by the row sum; dtype preserved)."""
    m = np.max(x, axis=-1, keepdims=True)
    e = np.exp(x - m)
    return e / np.sum(e, axis=-1, keepdims=True)


def pred_and_conf(probs: np.ndarray):
    """evaluators/vl_evaluator.py:68 (argmax) and :83 (confidence gather)."""
    preds = np.argmax(probs, axis=1)
    confs = probs[np.arange(probs.shape[0]), preds]
    return preds, confs


def score_chain(img, txt, class_confidence, logit_scale: float, chunk: int = 4096, threads=None):
    """features -> (pred, conf, top-2 logit gap) through the reference composition
    contraction -> DAC.predict -> softmax -> argmax/gather, row-chunked so that 49k-class
    vocabularies fit in host memory.  `class_confidence=None` skips DAC (fp32 kept)."""
    N = img.shape[0]
    preds = np.empty(N, np.int64)
    confs = np.empty(N, np.float32)
    gaps = np.empty(N, np.float32)
    for lo in range(0, N, chunk):
        hi = min(N, lo + chunk)
        lg = logits_fp32(img[lo:hi], txt, logit_scale, threads)
        if lg.shape[1] > 1:
            part = np.partition(lg, lg.shape[1] - 2, axis=1)
            gaps[lo:hi] = part[:, -1] - part[:, -2]
        else:
            gaps[lo:hi] = np.inf
        if class_confidence is not None:
            lg = dac_predict(lg, class_confidence)
        p, c = pred_and_conf(softmax_lastaxis(lg))
        preds[lo:hi], confs[lo:hi] = p, c
    return preds, confs, gaps


# --------------------------------------------------------------------------------------
# a-10 .. a-12  metrics
# --------------------------------------------------------------------------------------
def ece(conf, pred, gt, n_bins: int = 10) -> float:
    """tools/metrics.py:90-130.  Half-open digitize bins for the means (:104-120), closed
    last bin from np.histogram for the weights (:127)."""
    conf, pred, gt = np.asarray(conf), np.asarray(pred), np.asarray(gt)
    edges = np.linspace(0, 1, n_bins + 1)
    which = np.digitize(conf, edges) - 1
    acc = np.zeros(n_bins)
    mean_conf = np.zeros(n_bins)
    for b in range(n_bins):
        sel = which == b
        if sel.any():
            acc[b] = np.mean(gt[sel] == pred[sel])
            mean_conf[b] = np.mean(conf[sel])
    w = np.histogram(conf, edges)[0] / len(conf)
    return float(np.sum(w * np.abs(mean_conf - acc)))


def _grouped_gap(bin_id, conf, correct):
    """sum/max building block of MCE / AdaptiveECE / PIECE: per non-empty group
    |mean(correct) - mean(conf)| * count / N   (tools/metrics.py:203-206, :231-234)."""
    n = len(conf)
    out = []
    for b in np.unique(bin_id):
        sel = bin_id == b
        # pandas groupby().mean() Kahan-sums in the column's own dtype and divides by the count
        # in that dtype; emulated as "correctly rounded sum, then one division" (differs from a
        # true float32 Kahan loop by at most 1 ulp of the mean, ~6e-8, hence the pin tolerance)
        ty = conf.dtype.type
        mean_conf = ty(ty(np.sum(conf[sel].astype(np.float64))) / ty(sel.sum()))
        out.append(abs(np.mean(correct[sel].astype(np.float64)) - mean_conf) * sel.sum() / n)
    return np.array(out)


def mce(conf, pred, gt, n_bins: int = 10) -> float:
    """tools/metrics.py:181-208: inner edges only (conf==1.0 falls in the last bin) and the
    max of the *count-weighted* gaps."""
    conf, pred, gt = np.asarray(conf), np.asarray(pred), np.asarray(gt)
    inner = np.linspace(0, 1, n_bins + 1)[1:-1]
    which = np.digitize(conf, inner)
    return float(_grouped_gap(which, conf, (pred == gt)).max())


def quantile_edges(x: np.ndarray, n_bins: int, method: str = "averaged_inverted_cdf") -> np.ndarray:
    """Bin edges of sklearn 1.9 KBinsDiscretizer(strategy='quantile') as used at
    tools/metrics.py:228 (third-party: scikit-learn, unpinned in the reference's
    requirements; behaviour restated for the installed 1.9.0): percentiles at
    linspace(0,100,n+1), then edges closer than 1e-8 to their predecessor are dropped.
    (No subsampling here: sklearn subsamples to 200k rows with an unseeded RNG above that.)"""
    x = np.asarray(x)                      # percentiles are taken in the column's own dtype
    q = np.linspace(0, 100, n_bins + 1)
    edges = np.asarray(np.percentile(x, q, method=method), dtype=np.float64)
    keep = np.ediff1d(edges, to_begin=np.inf) > 1e-8
    return edges[keep]


def quantile_bin_ids(x: np.ndarray, n_bins: int, method: str = "averaged_inverted_cdf") -> np.ndarray:
    x = np.asarray(x)
    if x.min() == x.max():
        return np.zeros(len(x), np.int64)            # constant feature -> single bin
    edges = quantile_edges(x, n_bins, method)
    return np.searchsorted(edges[1:-1], x.astype(np.float64), side="right")


def adaptive_ece(conf, pred, gt, n_bins: int = 10, method: str = "averaged_inverted_cdf") -> float:
    """tools/metrics.py:212-236 (equal-mass bins, sum of count-weighted gaps)."""
    conf, pred, gt = np.asarray(conf), np.asarray(pred), np.asarray(gt)
    which = quantile_bin_ids(conf, n_bins, method)
    return float(_grouped_gap(which, conf, (pred == gt)).sum())


def uniform_bin_ids(x, n_bins: int) -> np.ndarray:
    """KBinsDiscretizer(strategy='uniform', encode='ordinal') on one column (scikit-learn 1.9
    preprocessing/_discretization.py): edges = np.linspace(min, max, n_bins + 1) in the column's own dtype,
    bin = searchsorted(edges[1:-1], x, side='right'); a constant column collapses to one bin."""
    x = np.asarray(x)
    if x.min() == x.max():
        return np.zeros(len(x), np.int64)
    edges = np.linspace(x.min(), x.max(), n_bins + 1)
    return np.searchsorted(edges[1:-1], x, side="right")


def piece(conf, knndist, pred, gt, dist_bin_num: int = 10, conf_bin_num: int = 10,
          method: str = "averaged_inverted_cdf", knn_strategy: str = "quantile") -> float:
    """tools/metrics.py:132-178: 2-D groups (quantile - the default - or uniform bin of knndist) x
    (uniform inner-edge bin of conf)."""
    conf, pred, gt, knndist = map(np.asarray, (conf, pred, gt, knndist))
    kb = quantile_bin_ids(knndist, dist_bin_num, method) if knn_strategy == "quantile" else uniform_bin_ids(knndist, dist_bin_num)
    cb = np.digitize(conf, np.linspace(0, 1, conf_bin_num + 1)[1:-1])
    return float(_grouped_gap(kb * (conf_bin_num + 1) + cb, conf, (pred == gt)).sum())


# --------------------------------------------------------------------------------------
# the (n+1)-bin table the CUDA path emits, restated on CPU (SURVEY.md 8a cross-check)
# --------------------------------------------------------------------------------------
FX_SHIFT = 40


def bin_table(conf, pred, gt, thresholds) -> np.ndarray:
    """table[b] = (count, n_correct, sum of round(conf * 2^40)) with b = #(thresholds <= conf)."""
    conf = np.asarray(conf)
    thr = np.asarray(thresholds, dtype=np.float64)
    which = np.searchsorted(thr, conf.astype(np.float64), side="right")
    correct = (np.asarray(pred) == np.asarray(gt))
    fx = np.rint(conf.astype(np.float64) * float(1 << FX_SHIFT)).astype(np.uint64)
    tab = np.zeros((len(thr) + 1, 3), np.uint64)
    for b in range(len(thr) + 1):
        sel = which == b
        tab[b] = (sel.sum(), correct[sel].sum(), fx[sel].sum(dtype=np.uint64))
    return tab


# --------------------------------------------------------------------------------------
# a-4  temperature-scaling objective
# --------------------------------------------------------------------------------------
def ts_loss_and_grad(img, txt, labels, log_scale: float):
    """loss = F.cross_entropy(exp(t) * img @ txt.T, label) and dloss/dt by autograd
    (trainers/calibration/tempscaling.py:31-41, :53-56, :155-160), float64 on CPU."""
    import torch
    a = torch.from_numpy(np.asarray(img, dtype=np.float64))
    b = torch.from_numpy(np.asarray(txt, dtype=np.float64))
    y = torch.from_numpy(np.asarray(labels, dtype=np.int64))
    t = torch.tensor(float(log_scale), dtype=torch.float64, requires_grad=True)
    loss = torch.nn.functional.cross_entropy(t.exp() * a @ b.t(), y)
    loss.backward()
    return float(loss.detach()), float(t.grad)


# --------------------------------------------------------------------------------------
# f-1  image-proximity kNN distances
# --------------------------------------------------------------------------------------
def knn_dists(ref_rows, queries, k: int, drop_self: bool = False) -> np.ndarray:
    """trainers/calibration/proximity.py:19-46 (get_knn_dists) and, with drop_self,
    :49-70 (get_val_image_knn_dists: k+1 nearest, first dropped), fp32."""
    import torch
    r = torch.from_numpy(np.asarray(ref_rows, dtype=np.float32))
    q = torch.from_numpy(np.asarray(queries, dtype=np.float32))
    out = []
    kk = k + 1 if drop_self else k
    for f in q:
        d = torch.norm(r - f, dim=1)
        top, _ = torch.topk(d, k=kk, largest=False)
        out.append(top[1:].numpy() if drop_self else top.numpy())
    return np.array(out)


# --------------------------------------------------------------------------------------
# f-4  density-ratio (proximity-informed) calibration
# --------------------------------------------------------------------------------------
class KDEMultivariateCC:
    """Restatement of `statsmodels.api.nonparametric.KDEMultivariate(data=[dep, indep], var_type='cc',
    bw='normal_reference')` as the reference calls it (trainers/calibration/density_ratio_calibration.py:66,
    :70) and of its `.pdf(data_predict)` (:104-105).

    statsmodels is an UNPINNED third-party dependency (reference requirements.txt:7, no version) and is NOT
    installed in the build container, so this follows its published algorithm (statsmodels 0.14,
    nonparametric/_kernel_base.py `GenericKDE._normal_reference`, `gpke`; kernels.py `gaussian`):
        bw_j   = 1.06 * std_j(ddof=0) * nobs ** (-1 / (4 + k_vars))
        pdf(x) = 1/nobs * sum_i prod_j [ exp(-(X_ij - x_j)^2 / (2 bw_j^2)) / sqrt(2 pi) ] / prod_j bw_j
    all in float64.  PARITY UNPINNED against statsmodels itself; cross-checked against
    sklearn.neighbors.KernelDensity on bandwidth-scaled data in oracle/make_golden.py."""

    def __init__(self, data, var_type="cc", bw="normal_reference"):
        assert var_type == "cc" and bw == "normal_reference"
        dat = np.asarray([np.asarray(v, dtype=np.float64) for v in data])     # [k_vars, nobs]
        self.k_vars = dat.shape[0]
        self.data = dat.T.reshape(-1, self.k_vars)                             # _adjust_shape -> [nobs, k_vars]
        self.nobs = self.data.shape[0]
        self.bw = 1.06 * np.std(self.data, axis=0) * self.nobs ** (-1.0 / (4 + self.k_vars))

    def pdf(self, data_predict):
        pts = np.asarray(data_predict, dtype=np.float64).reshape(-1, self.k_vars)
        out = np.empty(len(pts), np.float64)
        for i, x in enumerate(pts):
            kval = (1.0 / np.sqrt(2 * np.pi)) * np.exp(-(self.data - x) ** 2 / (self.bw ** 2 * 2.0))
            out[i] = (kval.prod(axis=1) / np.prod(self.bw)).sum(axis=0) / self.nobs
        return out


def density_ratio_fit(probs, preds, true, proximity):
    """trainers/calibration/density_ratio_calibration.py:35-72 (DensityRatioCalibration.fit): KDE of
    (confidence, proximity) over the correctly classified and over the misclassified validation samples,
    and the false/true count ratio."""
    probs = np.asarray(probs)
    assert np.all(probs >= 0) and np.all(probs <= 1)
    confs = np.max(probs, axis=-1)
    correct = np.asarray(preds) == np.asarray(true)
    proximity = np.asarray(proximity)
    dens_true = KDEMultivariateCC([confs[correct], proximity[correct]])
    dens_false = KDEMultivariateCC([confs[~correct], proximity[~correct]])
    ratio = (~correct).sum() / correct.sum()
    return dens_true, dens_false, ratio


def density_ratio_predict(state, probs, proximities):
    """trainers/calibration/density_ratio_calibration.py:78-117 (DensityRatioCalibration.predict): Bayes
    posterior of `correct` given (confidence, proximity); the predicted class gets that value, the other
    classes are rescaled to sum to one minus it.  Returns (probs_out [N, C] float64, conf_calibrated [N])."""
    dens_true, dens_false, ratio = state
    probs = np.asarray(probs)
    preds = np.argmax(probs, axis=-1)
    confs = np.max(probs, axis=-1)
    data = np.array([confs, proximities]).T
    t = dens_true.pdf(data)
    f = dens_false.pdf(data)
    cal = t / np.maximum(t + f * ratio, 1e-10)
    mask = np.ones(probs.shape, dtype=bool)
    mask[np.arange(probs.shape[0]), preds] = False
    rest = probs * mask
    out = rest * ((1 - cal) / rest.sum(axis=-1))[:, np.newaxis]
    out[np.arange(probs.shape[0]), preds] = cal
    return out, cal


# --------------------------------------------------------------------------------------
# f-4  multi-class isotonic regression and its proximity-binned wrapper
# --------------------------------------------------------------------------------------
def multi_isotonic_fit_transform(logit, label):
    """trainers/calibration/multi_isotonic_regression.py:14-30 (MultiIsotonicRegression.fit_transform): softmax
    WITHOUT max shift of whatever is passed as `logit` (the caller passes probabilities), one-hot labels, scikit-learn
    IsotonicRegression(out_of_bounds='clip') on the flattened arrays, `+ 1e-9 * p`.  Returns (p_out, calibrator).
    scikit-learn is the reference's own dependency for this step, so the oracle calls it like the reference does."""
    from sklearn.isotonic import IsotonicRegression
    logit = np.asarray(logit)
    label = np.asarray(label)
    n_classes = logit.shape[1]
    if label.ndim == 1:
        onehot = np.zeros((len(label), n_classes))
        onehot[np.arange(len(label)), label] = 1          # == label_binarize(label, classes=arange(C)) for C > 2
        label = onehot
    p = np.exp(logit) / np.sum(np.exp(logit), 1)[:, None]
    calibrator = IsotonicRegression(out_of_bounds="clip")
    y_ = calibrator.fit_transform(p.flatten(), label.flatten())
    return y_.reshape(logit.shape) + 1e-9 * p, calibrator


def multi_isotonic_transform(calibrator, logit):
    """multi_isotonic_regression.py:32-35."""
    logit = np.asarray(logit)
    p = np.exp(logit) / np.sum(np.exp(logit), 1)[:, None]
    return calibrator.predict(p.flatten()).reshape(logit.shape) + 1e-9 * p


def bin_mean_shift_fit_transform(logit, proximity, label, proximity_bin: int = 5, bin_strategy: str = "quantile"):
    """trainers/calibration/multi_proximity_isotonic.py:198-229 (BinMeanShift.fit_transform with the
    'multi_isotonic_regression' method, normalize_conf=False): proximity bins by np.percentile edges (:158-160) or
    uniform edges (:174-176), one MultiIsotonicRegression per bin.  Returns (probs, (edges, calibrators))."""
    logit, proximity, label = np.asarray(logit), np.asarray(proximity), np.asarray(label)
    if bin_strategy == "quantile":
        edges = np.asarray(np.percentile(proximity, np.linspace(0, 100, proximity_bin + 1)))
    else:
        edges = np.linspace(proximity.min(), proximity.max(), proximity_bin + 1)
    bin_no = np.searchsorted(edges[1:-1], proximity, side="right")
    out = np.empty(logit.shape, np.float64)
    cals = []
    for b in range(proximity_bin):
        idx = np.nonzero(bin_no == b)[0]
        res, cal = multi_isotonic_fit_transform(logit[idx], label[idx])
        out[idx] = res
        cals.append(cal)
    return out, (edges, cals)


def bin_mean_shift_transform(state, logit, proximity):
    """multi_proximity_isotonic.py:231-247."""
    edges, cals = state
    logit, proximity = np.asarray(logit), np.asarray(proximity)
    bin_no = np.searchsorted(edges[1:-1], proximity, side="right")
    out = np.empty(logit.shape, np.float64)
    for b in range(len(cals)):
        idx = np.nonzero(bin_no == b)[0]
        if len(idx):
            out[idx] = multi_isotonic_transform(cals[b], logit[idx])
    return out


class _NetcalOneVsAllCC:
    """netcal.AbstractCalibration for multi-class input, `detection=False` (published algorithm of netcal 1.3:
    `_create_one_vs_all_models` + `_calibrate_multiclass`): a binary sub-model per class that occurs in y, fitted on
    (X[:, j], y == j); transform puts sub-model j's output into column j (classes without a sub-model stay 0) and
    divides each row by its sum unless `independent_probabilities`.  A 1-D X is the binary problem itself.  netcal is
    called by the reference at trainers/calibration/vl_calibrator.py:125-131, :137-143 and
    multi_proximity_isotonic.py:219-221, :241-243.  PARITY WITH NETCAL UNPINNED (not installed)."""

    def __init__(self, independent_probabilities=False):
        self.independent_probabilities = independent_probabilities
        self.models = None
        self.binary = False

    def fit(self, X, y):
        X = np.asarray(X, dtype=np.float64)
        y = np.asarray(y)
        if y.ndim == 2:
            y = np.argmax(y, axis=1)
        self.binary = X.ndim == 1
        if self.binary:
            self.models = [self._fit_binary(X, (y == 1).astype(np.float64))]
            return self
        self.models = []
        for j in range(X.shape[1]):
            hit = y == j
            self.models.append(self._fit_binary(X[:, j], hit.astype(np.float64)) if hit.any() else None)
        return self

    def transform(self, X):
        X = np.asarray(X, dtype=np.float64)
        if self.binary:
            return self._apply_binary(self.models[0], X)
        out = np.zeros(X.shape, np.float64)
        for j, m in enumerate(self.models):
            if m is not None:
                out[:, j] = self._apply_binary(m, X[:, j])
        if not self.independent_probabilities:
            with np.errstate(invalid="ignore", divide="ignore"):
                out = out / np.sum(out, axis=1, keepdims=True)
        return out

    def fit_transform(self, X, y):
        return self.fit(X, y).transform(X)


class NetcalHistogramBinningCC(_NetcalOneVsAllCC):
    """netcal.binning.HistogramBinning(bins, equal_intervals=True): bin bounds np.linspace(0, 1, bins + 1), bin map =
    scipy.stats.binned_statistic_dd(X, y, 'mean') (last bin closed), empty bins filled with the bin centre; transform
    looks the bin up with np.digitize - 1 clipped to the valid range."""

    def __init__(self, bins=10, equal_intervals=True, detection=False, independent_probabilities=False):
        super().__init__(independent_probabilities)
        assert equal_intervals and not detection
        self.bins = bins

    def _fit_binary(self, x, y):
        from scipy.stats import binned_statistic_dd
        edges = np.linspace(0.0, 1.0, self.bins + 1)
        mean, _, _ = binned_statistic_dd(x.reshape(-1, 1), y, statistic="mean", bins=[edges])
        centres = (edges[1:] + edges[:-1]) * 0.5
        return edges, np.where(np.isnan(mean), centres, mean)

    def _apply_binary(self, model, x):
        edges, bin_map = model
        idx = np.clip(np.digitize(x, edges) - 1, 0, self.bins - 1)
        return bin_map[idx]


class NetcalIsotonicRegressionCC(_NetcalOneVsAllCC):
    """netcal.binning.IsotonicRegression: sklearn.isotonic.IsotonicRegression(increasing=True, out_of_bounds='clip')
    per binary problem."""

    def __init__(self, detection=False, independent_probabilities=False):
        super().__init__(independent_probabilities)
        assert not detection

    def _fit_binary(self, x, y):
        from sklearn.isotonic import IsotonicRegression
        return IsotonicRegression(increasing=True, out_of_bounds="clip").fit(x, y)

    def _apply_binary(self, model, x):
        return model.transform(x)
