"""Generate tests/golden/*.npz by running the REAL reference functions (imported from
/root/reference, which only exists in the build container) on the seeded synthetic cases,
and pin oracle/cpu_oracle.py against them while doing so.

    python oracle/make_golden.py            # writes tests/golden/, asserts oracle == reference

Reference entry points executed unmodified:
  trainers/calibration/distanse_aware_calibration.py  DistanseAwareCalibration.fit/.predict
      (.cuda() patched to identity: the container has no GPU)
  tools/metrics.py   ECE, MCE, AdaptiveECE, PIECE
  trainers/calibration/proximity.py   get_knn_dists, get_val_image_knn_dists (.to('cuda') patched)
  trainers/calibration/density_ratio_calibration.py  DensityRatioCalibration.fit/.predict (statsmodels stubbed
      with the oracle's restatement of KDEMultivariate - the package is not installed)
  trainers/calibration/multi_isotonic_regression.py MultiIsotonicRegression, multi_proximity_isotonic.py BinMeanShift
and the glue the reference performs inline: `(s*img)@txt.T` in fp32 torch
(trainers/classification/zsclip.py:97-102), scipy.special.softmax(axis=-1)
(trainers/calibration/vl_calibrator.py:91), argmax + gather (evaluators/vl_evaluator.py:68,:83).
"""
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("CCAL_REFERENCE", "/root/reference")
sys.path.insert(1, REF)

import torch  # noqa: E402
from scipy.special import softmax  # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self          # no GPU in the build container
_orig_to = torch.Tensor.to
torch.Tensor.to = lambda self, *a, **k: self if (a and a[0] == "cuda") else _orig_to(self, *a, **k)

import tools.metrics as ref_metrics  # noqa: E402
from trainers.calibration.distanse_aware_calibration import DistanseAwareCalibration  # noqa: E402
import trainers.calibration.proximity as ref_prox  # noqa: E402

from clip_calibration_b200 import synth  # noqa: E402
from oracle import cpu_oracle as orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
warnings.filterwarnings("ignore")


def ref_chain(case, cc, chunk=4096):
    N = case.img.shape[0]
    preds = np.empty(N, np.int64)
    confs = np.empty(N, np.float32)
    dac = DistanseAwareCalibration()
    dac.class_confidence = cc
    img = torch.from_numpy(case.img)
    txt = torch.from_numpy(case.txt_tuned)
    for lo in range(0, N, chunk):
        hi = min(N, lo + chunk)
        logits = (torch.tensor(case.logit_scale) * img[lo:hi] @ txt.t()).numpy()
        # the evaluator round-trips logits through Python floats -> float64 (base_learner.py:94)
        scaled = dac.predict(logits.astype(np.float64)) if cc is not None else logits
        probs = softmax(scaled, axis=-1)
        p = np.argmax(probs, axis=1)
        preds[lo:hi] = p
        confs[lo:hi] = probs[np.arange(hi - lo), p]
    return preds, confs


def run_case(name, n_override=None, ks=(5,), fit_classes=None, seed=0):
    t0 = time.time()
    case = synth.make_config(name, seed=seed, n_override=n_override)
    out = {"seed": seed, "N": case.img.shape[0], "C": case.txt_zs.shape[0], "n_base": case.n_base,
           "signal": case.signal, "logit_scale": case.logit_scale,
           "img_checksum": float(case.img.astype(np.float64).sum()),
           "txt_checksum": float(case.txt_tuned.astype(np.float64).sum())}
    for k in ks:
        dac = DistanseAwareCalibration()
        dac.fit(case.base_zs, case.txt_zs, case.base_tuned, case.txt_tuned, k)
        cc = dac.class_confidence
        out[f"cc_k{k}"] = cc
        # pin the oracle's fit (on a class subset when the vocabulary is large)
        sel = np.arange(len(cc)) if fit_classes is None else np.linspace(0, len(cc) - 1, fit_classes).astype(int)
        occ, iz, it, dz, dt = orc.dac_fit(case.base_zs, case.txt_zs[sel], case.base_tuned, case.txt_tuned[sel], k)
        assert np.array_equal(occ, cc[sel]), f"{name}: oracle dac_fit != reference"
        out[f"fit_sel_k{k}"] = sel
        out[f"knn_idx_zs_k{k}"] = iz.astype(np.int32)
        out[f"knn_idx_tuned_k{k}"] = it.astype(np.int32)
        out[f"knn_dist_tuned_k{k}"] = dt
    k = ks[len(ks) // 2] if len(ks) > 1 else ks[0]
    cc = out[f"cc_k{k}"]
    for tag, ccx in (("dac", cc), ("nodac", None)):
        pred, conf = ref_chain(case, ccx)
        opred, oconf, gap = orc.score_chain(case.img, case.txt_tuned, ccx, case.logit_scale)
        assert np.array_equal(pred, opred), f"{name}/{tag}: oracle pred != reference"
        assert np.allclose(conf, oconf, rtol=2e-6, atol=0), f"{name}/{tag}: oracle conf != reference"
        res = {"pred": pred.astype(np.int32), "conf": conf, "gap": gap}
        for nb in (10, 15):
            e = ref_metrics.ECE(conf, pred, case.labels, nb)
            m = ref_metrics.MCE(conf, pred, case.labels, nb)
            a = ref_metrics.AdaptiveECE(conf, pred, case.labels, nb)
            assert abs(orc.ece(conf, pred, case.labels, nb) - e) < 1e-12
            assert abs(orc.mce(conf, pred, case.labels, nb) - m) < 3e-8
            if len(conf) <= 200000:
                assert abs(orc.adaptive_ece(conf, pred, case.labels, nb) - a) < 3e-8, (name, tag, nb)
            res[f"ece{nb}"], res[f"mce{nb}"], res[f"ace{nb}"] = float(e), float(m), float(a)
        res["acc"] = float(np.mean(pred == case.labels))
        res["mean_conf"] = float(np.mean(conf))
        res["counts10"] = np.histogram(conf, np.linspace(0, 1, 11))[0]
        for kk, v in res.items():
            out[f"{tag}_{kk}"] = v
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(f"{name}: N={out['N']} C={out['C']} acc={out['dac_acc']:.6f} ece10={out['dac_ece10']:.8f} "
          f"mce10={out['dac_mce10']:.8f} ace10={out['dac_ace10']:.8f} nodac_ece10={out['nodac_ece10']:.8f} "
          f"({time.time()-t0:.1f}s)")


def metric_edge_cases():
    """Small hand-built confidence vectors that hit the binning corner cases
    (SURVEY.md Appendix A.1-A.3): conf == 1.0, conf exactly on float64 edges, empty bins,
    float64 confidences, a single sample, all-equal confidences."""
    rng = np.random.default_rng(7)
    cases = {}
    e10 = np.linspace(0, 1, 11)
    cases["ones_and_half"] = (np.array([1.0, 0.5], np.float32), np.array([0, 0]), np.array([0, 0]))
    on_edges = np.concatenate([e10.astype(np.float32), np.nextafter(e10.astype(np.float32), np.float32(0)),
                               np.nextafter(e10.astype(np.float32), np.float32(2))])
    on_edges = np.clip(on_edges, 0, 1).astype(np.float32)
    cases["on_edges"] = (on_edges, rng.integers(0, 3, on_edges.size), rng.integers(0, 3, on_edges.size))
    sat = np.where(rng.random(4000) < 0.4, 1.0, rng.random(4000)).astype(np.float32)
    cases["saturated"] = (sat, rng.integers(0, 4, 4000), rng.integers(0, 4, 4000))
    f64 = rng.random(3000)
    cases["float64_conf"] = (f64, rng.integers(0, 2, 3000), rng.integers(0, 2, 3000))
    cases["single"] = (np.array([0.3], np.float32), np.array([1]), np.array([1]))
    narrow = (0.55 + 0.1 * rng.random(2500)).astype(np.float32)
    cases["narrow"] = (narrow, rng.integers(0, 2, 2500), rng.integers(0, 2, 2500))
    dup = np.round(rng.random(5000) * 20) / 20
    cases["many_ties"] = (dup.astype(np.float32), rng.integers(0, 2, 5000), rng.integers(0, 2, 5000))
    out = {}
    for name, (conf, pred, gt) in cases.items():
        out[f"{name}_conf"], out[f"{name}_pred"], out[f"{name}_gt"] = conf, pred.astype(np.int64), gt.astype(np.int64)
        for nb in (10, 15):
            e = ref_metrics.ECE(conf, pred, gt, nb)
            m = ref_metrics.MCE(conf, pred, gt, nb)
            a = ref_metrics.AdaptiveECE(conf, pred, gt, nb)
            assert abs(orc.ece(conf, pred, gt, nb) - e) < 1e-12, name
            assert abs(orc.mce(conf, pred, gt, nb) - m) < 3e-8, name
            assert abs(orc.adaptive_ece(conf, pred, gt, nb) - a) < 3e-8, (name, nb, orc.adaptive_ece(conf, pred, gt, nb), a)
            out[f"{name}_ece{nb}"], out[f"{name}_mce{nb}"], out[f"{name}_ace{nb}"] = float(e), float(m), float(a)
    out["names"] = np.array(sorted(cases))
    np.savez_compressed(os.path.join(OUT, "metric_edge_cases.npz"), **out)
    print("metric_edge_cases:", ", ".join(sorted(cases)))


def proximity_and_piece():
    """trainers/calibration/proximity.py + tools/metrics.py PIECE on a small case."""
    case = synth.make_case("prox", 600, 40, 20, 512, 5, 0.3, seed=3)
    val = synth.make_case("proxval", 300, 40, 20, 512, 5, 0.3, seed=4).img
    kd = ref_prox.get_knn_dists(val, case.img, 5)
    kd_self = ref_prox.get_val_image_knn_dists(val, 5)
    assert np.allclose(orc.knn_dists(val, case.img, 5), kd, rtol=0, atol=0)
    assert np.allclose(orc.knn_dists(val, val, 5, drop_self=True), kd_self, rtol=0, atol=0)
    pred, conf, _ = orc.score_chain(case.img, case.txt_tuned, None, case.logit_scale)
    prox = np.exp(-np.mean(kd, axis=-1))      # base_learner.py:137
    out = {"knn": kd, "knn_self": kd_self, "pred": pred, "conf": conf, "labels": case.labels}
    for nb in (10, 5):
        p = ref_metrics.PIECE(conf, prox, pred, case.labels, nb, 10)
        assert abs(orc.piece(conf, prox, pred, case.labels, nb, 10) - p) < 3e-8
        out[f"piece{nb}"] = float(p)
        pu = ref_metrics.PIECE(conf, prox, pred, case.labels, nb, 10, knn_strategy="uniform")      # tools/metrics.py:152
        assert abs(orc.piece(conf, prox, pred, case.labels, nb, 10, knn_strategy="uniform") - pu) < 3e-8
        out[f"piece{nb}_uniform"] = float(pu)
    np.savez_compressed(os.path.join(OUT, "proximity_piece.npz"), **out)
    print("proximity_piece: piece10=%.8f" % out["piece10"])


def density_ratio():
    """trainers/calibration/density_ratio_calibration.py DensityRatioCalibration.fit/.predict, run unmodified.
    statsmodels is not installed here: `statsmodels.api` is stubbed so that `sm.nonparametric.KDEMultivariate`
    resolves to the oracle's restatement of it (published algorithm) - that pins the REFERENCE's arithmetic around
    the KDE; the KDE restatement itself is cross-checked against sklearn.neighbors.KernelDensity below."""
    import types
    stub = types.ModuleType("statsmodels.api")
    stub.nonparametric = types.SimpleNamespace(KDEMultivariate=orc.KDEMultivariateCC)
    pkg = types.ModuleType("statsmodels")
    pkg.api = stub
    sys.modules.setdefault("statsmodels", pkg)
    sys.modules.setdefault("statsmodels.api", stub)
    from trainers.calibration.density_ratio_calibration import DensityRatioCalibration
    from sklearn.neighbors import KernelDensity

    both = synth.make_case("dr", 1600, 30, 15, 512, 5, 0.22, seed=11)      # one task; rows split into val / test
    n_val = 1000
    logits = (torch.tensor(both.logit_scale) * torch.from_numpy(both.img) @ torch.from_numpy(both.txt_tuned).t()).numpy()
    probs = softmax(logits.astype(np.float64) * 0.5, axis=1).astype(np.float32)
    val_probs32, test_probs32 = probs[:n_val], probs[n_val:]
    val_labels, test_labels = both.labels[:n_val], both.labels[n_val:]
    val_img, test_img = both.img[:n_val], both.img[n_val:]
    val_prox = np.exp(-np.mean(ref_prox.get_val_image_knn_dists(val_img, 10), axis=-1))      # vl_calibrator.py:68
    test_prox = np.exp(-np.mean(ref_prox.get_knn_dists(val_img, test_img, 10), axis=-1))     # base_learner.py:137
    val_preds = np.argmax(val_probs32, axis=1)
    out = {"val_probs": val_probs32, "val_preds": val_preds, "val_labels": val_labels, "val_prox": val_prox,
           "test_probs": test_probs32, "test_prox": test_prox, "test_labels": test_labels}
    for tag, vp, tp in (("f32", val_probs32, test_probs32), ("f64", val_probs32.astype(np.float64), test_probs32.astype(np.float64))):
        ref = DensityRatioCalibration()
        ref.fit(vp, val_preds, val_labels, val_prox)
        got = ref.predict(tp.copy(), test_prox)
        state = orc.density_ratio_fit(vp, val_preds, val_labels, val_prox)
        mine, cal = orc.density_ratio_predict(state, tp, test_prox)
        assert np.array_equal(got, mine), "oracle density_ratio_predict != reference"
        assert abs(state[2] - ref.false_true_ratio) == 0
        if tag == "f32":
            out["f32_probs_out"] = got
        out[f"{tag}_conf_cal"] = cal
        out[f"{tag}_bw_true"], out[f"{tag}_bw_false"], out[f"{tag}_ratio"] = state[0].bw, state[1].bw, float(state[2])
        # independent check of the KDE restatement: isotropic Gaussian KDE on bandwidth-scaled coordinates
        for dens in state[:2]:
            kd = KernelDensity(kernel="gaussian", bandwidth=1.0).fit(dens.data / dens.bw)
            pts = np.array([np.max(tp, axis=-1), test_prox]).T
            want = np.exp(kd.score_samples(pts / dens.bw)) / np.prod(dens.bw)
            assert np.allclose(dens.pdf(pts), want, rtol=1e-8, atol=1e-12 * want.max()), "KDE restatement != sklearn KernelDensity"
    print("density_ratio: n_val=%d (true %d) n_test=%d  mean conf %.6f -> %.6f  acc %.6f" % (
        len(val_preds), int((val_preds == val_labels).sum()), len(test_labels), float(np.max(test_probs32, 1).mean()),
        float(out["f64_conf_cal"].mean()), float((np.argmax(test_probs32, 1) == test_labels).mean())))
    np.savez_compressed(os.path.join(OUT, "density_ratio.npz"), **out)


def isotonic():
    """trainers/calibration/multi_isotonic_regression.py MultiIsotonicRegression and
    trainers/calibration/multi_proximity_isotonic.py BinMeanShift('multi_isotonic_regression', ..., 'quantile',
    proximity_bin=5) - the calls of vl_calibrator.py:133-134, :146-148 - run unmodified on the density-ratio case's
    probabilities (the caller passes probabilities where the class says `logit`)."""
    from trainers.calibration.multi_isotonic_regression import MultiIsotonicRegression
    from trainers.calibration.multi_proximity_isotonic import BinMeanShift
    g = np.load(os.path.join(OUT, "density_ratio.npz"))
    vp, tp = g["val_probs"].astype(np.float64), g["test_probs"].astype(np.float64)
    vl, vprox, tprox = g["val_labels"], g["val_prox"], g["test_prox"]
    ref = MultiIsotonicRegression()
    val_out = ref.fit_transform(vp, vl)
    test_out = ref.transform(tp)
    mine_val, cal = orc.multi_isotonic_fit_transform(vp, vl)
    assert np.array_equal(mine_val, val_out) and np.array_equal(orc.multi_isotonic_transform(cal, tp), test_out)
    rv, rt = np.arange(0, len(vp), 5), np.arange(0, len(tp), 4)        # stored row subsets keep the fixture small
    out = {"rows_val": rv, "rows_test": rt, "val_out": val_out[rv], "test_out": test_out[rt],
           "x_thresholds": ref.calibrator.X_thresholds_, "y_thresholds": ref.calibrator.y_thresholds_}
    for strategy in ("quantile", "uniform"):
        bms = BinMeanShift("multi_isotonic_regression", MultiIsotonicRegression, bin_strategy=strategy,
                           normalize_conf=False, proximity_bin=5)
        b_val = bms.fit_transform(vp, vprox, vl)
        b_test = bms.transform(tp, tprox)
        m_val, state = orc.bin_mean_shift_fit_transform(vp, vprox, vl, 5, strategy)
        assert np.array_equal(m_val, b_val) and np.array_equal(orc.bin_mean_shift_transform(state, tp, tprox), b_test)
        out[f"bms_{strategy}_edges"] = np.asarray(bms.bin_edges, dtype=np.float64)
        out[f"bms_{strategy}_val_out"], out[f"bms_{strategy}_test_out"] = b_val[rv], b_test[rt]
    print("isotonic: %d knots; BinMeanShift quantile edges %s" % (len(out["x_thresholds"]), np.round(out["bms_quantile_edges"], 5)))
    np.savez_compressed(os.path.join(OUT, "isotonic.npz"), **out)


def netcal_binning():
    """The two netcal calibrators `VLCalibration` can build (vl_calibrator.py:125-131 under BinMeanShift, :137-143
    plain).  netcal is not installed: the oracle's restatement of its published one-vs-all scheme stands in for
    `netcal.binning.{HistogramBinning, IsotonicRegression}` (PARITY WITH NETCAL UNPINNED) while the reference's own
    BinMeanShift (multi_proximity_isotonic.py:130-247: exp-normalisation of its input for these two methods, proximity
    bins, per-bin fit_transform / transform, re-ordering) runs unmodified around it."""
    from trainers.calibration.multi_proximity_isotonic import BinMeanShift
    g = np.load(os.path.join(OUT, "density_ratio.npz"))
    vp, tp = g["val_probs"].astype(np.float64), g["test_probs"].astype(np.float64)   # as in isotonic()
    vl, vprox, tprox = g["val_labels"], g["val_prox"], g["test_prox"]
    rv, rt = np.arange(0, len(vp), 5), np.arange(0, len(tp), 4)
    out = {"rows_val": rv, "rows_test": rt}
    for name, cls, kw in (("histogram_binning", orc.NetcalHistogramBinningCC, {"bins": 10}),
                          ("isotonic_regression", orc.NetcalIsotonicRegressionCC, {})):
        plain = cls(**kw)
        plain.fit(vp, vl)                                               # :138-143
        out[f"{name}_val_out"], out[f"{name}_test_out"] = plain.transform(vp)[rv], plain.transform(tp)[rt]
        bms = BinMeanShift(name, cls, bin_strategy="quantile", normalize_conf=False, proximity_bin=5, **kw)
        b_val = bms.fit_transform(vp, vprox, vl)                        # :126-131
        b_test = bms.transform(tp, tprox)
        out[f"bms_{name}_edges"] = np.asarray(bms.bin_edges, dtype=np.float64)
        out[f"bms_{name}_val_out"], out[f"bms_{name}_test_out"] = b_val[rv], b_test[rt]
        print("netcal_binning %s: plain test row sums %.6f..%.6f, BinMeanShift NaN rows %d" % (
            name, np.nanmin(plain.transform(tp).sum(1)), np.nanmax(plain.transform(tp).sum(1)),
            int(np.isnan(b_test).any(axis=1).sum())))
    np.savez_compressed(os.path.join(OUT, "netcal_binning.npz"), **out)


def in21k_fit(fit_classes=256, seed=0):
    """DAC fit at BASELINE.json configs[4]'s own shape - 21,841 test classes x 10,000 base classes x 768-d - on a
    class subset (the reference loop costs ~25 ms per class here): the reference's class_confidence and the
    oracle's neighbour lists for `fit_classes` classes spread over the vocabulary."""
    t0 = time.time()
    N, C, B, D, k, a = synth.CONFIGS["in21k"]
    txt_zs, txt_tuned, _ = synth.make_text(C, D, seed)
    sel = np.linspace(0, C - 1, fit_classes).astype(int)
    dac = DistanseAwareCalibration()
    dac.fit(txt_zs[:B], txt_zs[sel], txt_tuned[:B], txt_tuned[sel], k)
    cc = np.asarray(dac.class_confidence)
    occ, iz, it, dz, dt = orc.dac_fit(txt_zs[:B], txt_zs[sel], txt_tuned[:B], txt_tuned[sel], k)
    assert np.array_equal(occ, cc), "in21k: oracle dac_fit != reference"
    np.savez_compressed(os.path.join(OUT, "in21k_fit.npz"), seed=seed, C=C, B=B, D=D, k=k, sel=sel, cc=cc,
                        knn_idx_zs=iz.astype(np.int32), knn_idx_tuned=it.astype(np.int32), knn_dist_tuned=dt,
                        txt_checksum=float(txt_tuned.astype(np.float64).sum()))
    print(f"in21k_fit: {fit_classes} of {C} classes x {B} base x {D}: cc in [{cc.min():.6f}, {cc.max():.6f}] ({time.time()-t0:.1f}s)")


def dac_float16():
    """The reference's DAC fit on FLOAT16 feature arrays - its default precision (train.py:152); numpy keeps float16
    through np.linalg.norm / np.sum / np.exp (distanse_aware_calibration.py:28-42).  Pins the opt-in
    arithmetic="input" mode (ccal_dac_fit_f16) and measures how far the float32 answer is from it."""
    t0 = time.time()
    out = {}
    cases = [("sun397_l14", None, (1, 5, 10)), ("openvocab", 256, (5,)), ("in21k", 48, (5,))]
    for name, fit_classes, ks in cases:
        N, C, B, D, _, a = synth.CONFIGS[name]
        txt_zs, txt_tuned, _ = synth.make_text(C, D, 0, rounding=synth.round_to_fp16)
        zs16, tu16 = txt_zs.astype(np.float16), txt_tuned.astype(np.float16)
        sel = np.arange(C) if fit_classes is None else np.linspace(0, C - 1, fit_classes).astype(int)
        out[f"{name}_sel"] = sel
        for k in ks:
            dac = DistanseAwareCalibration()
            dac.fit(zs16[:B], zs16[sel], tu16[:B], tu16[sel], k)
            cc16 = np.asarray(dac.class_confidence, dtype=np.float64)
            occ, iz, it, dz, dt = orc.dac_fit(zs16[:B], zs16[sel], tu16[:B], tu16[sel], k)
            assert np.array_equal(np.asarray(occ, np.float64), cc16), f"{name}: oracle float16 fit != reference"
            cc32 = np.asarray(orc.dac_fit(txt_zs[:B], txt_zs[sel], txt_tuned[:B], txt_tuned[sel], k)[0], np.float64)
            out[f"{name}_cc16_k{k}"] = cc16
            out[f"{name}_dist16_tuned_k{k}"] = np.asarray(dt, np.float32)
            out[f"{name}_idx16_tuned_k{k}"] = it.astype(np.int32)
            out[f"{name}_rel_f32_vs_f16_k{k}"] = float(np.max(np.abs(cc32 - cc16) / cc16))
            print(f"dac_float16 {name} k={k}: {len(sel)} classes, max rel |cc32 - cc16| = {out[f'{name}_rel_f32_vs_f16_k{k}']:.2e}")
    np.savez_compressed(os.path.join(OUT, "dac_float16.npz"), **out)
    print(f"dac_float16 written ({time.time()-t0:.1f}s)")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    if len(sys.argv) > 1:                      # python oracle/make_golden.py netcal_binning [...]: only those fixtures
        for fn in sys.argv[1:]:
            globals()[fn]()
        sys.exit(0)
    metric_edge_cases()
    proximity_and_piece()
    density_ratio()
    isotonic()
    netcal_binning()
    run_case("eurosat")
    run_case("sun397_l14", ks=(1, 5, 10))
    run_case("imagenet")
    run_case("openvocab", n_override=1024, fit_classes=256)
    in21k_fit()
    dac_float16()
    print("golden fixtures written to", OUT)
