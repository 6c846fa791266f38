"""CPU: the oracle restatement against the golden vectors produced by the real reference
functions (oracle/make_golden.py).  Integer / index results must be identical; float results
within the stated tolerances (MKL thread count and ISA may move last digits of the GEMM)."""
import numpy as np
import pytest

from oracle import cpu_oracle as orc


@pytest.mark.parametrize("name", ["eurosat", "sun397_l14", "imagenet", "openvocab"])
def test_score_chain_matches_reference(name, golden, synth_case):
    g, case = golden(name), synth_case(name)
    k = 5
    cc = g[f"cc_k{k}"]
    for tag, ccx in (("dac", cc), ("nodac", None)):
        pred, conf, gap = orc.score_chain(case.img, case.txt_tuned, ccx, case.logit_scale)
        ties = g[f"{tag}_gap"] <= 4e-5
        assert np.array_equal(pred[~ties], g[f"{tag}_pred"][~ties])
        np.testing.assert_allclose(conf[~ties], g[f"{tag}_conf"][~ties], rtol=2e-5)
        assert abs(orc.ece(conf, pred, case.labels, 10) - float(g[f"{tag}_ece10"])) < 1e-6


@pytest.mark.parametrize("name,ks", [("eurosat", (5,)), ("sun397_l14", (1, 5, 10)), ("imagenet", (5,)), ("openvocab", (5,))])
def test_dac_fit_matches_reference(name, ks, golden, synth_case):
    g, case = golden(name), synth_case(name)
    for k in ks:
        sel = g[f"fit_sel_k{k}"]
        if name == "imagenet":
            sel = sel[::8]
        cc, iz, it, dz, dt = orc.dac_fit(case.base_zs, case.txt_zs[sel], case.base_tuned, case.txt_tuned[sel], k)
        ref_sel = np.searchsorted(g[f"fit_sel_k{k}"], sel)
        assert np.array_equal(cc, g[f"cc_k{k}"][sel])
        assert np.array_equal(it, g[f"knn_idx_tuned_k{k}"][ref_sel])
        assert np.array_equal(iz, g[f"knn_idx_zs_k{k}"][ref_sel])
    # base classes are calibrated with multiplier exactly 1 (self distance 0 < 0.05)
    assert np.all(g[f"cc_k{ks[0]}"][: int(g["n_base"])] == 1.0)


def test_metric_edge_cases(golden):
    g = golden("metric_edge_cases")
    for name in g["names"]:
        conf, pred, gt = g[f"{name}_conf"], g[f"{name}_pred"], g[f"{name}_gt"]
        for nb in (10, 15):
            assert abs(orc.ece(conf, pred, gt, nb) - float(g[f"{name}_ece{nb}"])) < 1e-12, name
            assert abs(orc.mce(conf, pred, gt, nb) - float(g[f"{name}_mce{nb}"])) < 3e-8, name
            assert abs(orc.adaptive_ece(conf, pred, gt, nb) - float(g[f"{name}_ace{nb}"])) < 3e-8, name


def test_reference_quirk_conf_equal_one():
    # the 1.0 sample is in no bin mean but in the last bin's weight
    assert orc.ece(np.array([1.0, 0.5], np.float32), np.array([0, 0]), np.array([0, 0]), 10) == 0.25


def test_proximity_and_piece(golden):
    from clip_calibration_b200 import synth
    g = golden("proximity_piece")
    case = synth.make_case("prox", 600, 40, 20, 512, 5, 0.3, seed=3)
    val = synth.make_case("proxval", 300, 40, 20, 512, 5, 0.3, seed=4).img
    np.testing.assert_allclose(orc.knn_dists(val, case.img, 5), g["knn"], rtol=1e-6)
    np.testing.assert_allclose(orc.knn_dists(val, val, 5, drop_self=True), g["knn_self"], rtol=1e-6, atol=1e-7)
    prox = np.exp(-np.mean(g["knn"], axis=-1))
    for nb in (10, 5):
        assert abs(orc.piece(g["conf"], prox, g["pred"], g["labels"], nb, 10) - float(g[f"piece{nb}"])) < 3e-8
        assert abs(orc.piece(g["conf"], prox, g["pred"], g["labels"], nb, 10, knn_strategy="uniform")
                   - float(g[f"piece{nb}_uniform"])) < 3e-8


def test_ts_loss_grad_against_finite_difference():
    from clip_calibration_b200 import synth
    case = synth.make_case("ts", 200, 37, 19, 128, 5, 0.3, seed=5)
    t = 4.6052
    loss, grad = orc.ts_loss_and_grad(case.img, case.txt_tuned, case.labels, t)
    lp, _ = orc.ts_loss_and_grad(case.img, case.txt_tuned, case.labels, t + 1e-5)
    lm, _ = orc.ts_loss_and_grad(case.img, case.txt_tuned, case.labels, t - 1e-5)
    assert abs((lp - lm) / 2e-5 - grad) < 1e-6 * max(1.0, abs(grad))


def test_density_ratio_matches_reference(golden):
    """The fixture holds what the reference's DensityRatioCalibration.fit/.predict returned (oracle/make_golden.py)."""
    g = golden("density_ratio")
    state = orc.density_ratio_fit(g["val_probs"], g["val_preds"], g["val_labels"], g["val_prox"])
    np.testing.assert_allclose(state[0].bw, g["f32_bw_true"], rtol=1e-13)
    np.testing.assert_allclose(state[1].bw, g["f32_bw_false"], rtol=1e-13)
    assert state[2] == float(g["f32_ratio"])
    out, cal = orc.density_ratio_predict(state, g["test_probs"], g["test_prox"])
    np.testing.assert_allclose(cal, g["f32_conf_cal"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(out, g["f32_probs_out"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(out.sum(axis=1), 1.0, rtol=1e-6)
    # float64 probabilities (the no-DAC branch keeps float64): same calibrated confidences here because the
    # fixture's float64 case is the float32 matrix widened
    state64 = orc.density_ratio_fit(g["val_probs"].astype(np.float64), g["val_preds"], g["val_labels"], g["val_prox"])
    _, cal64 = orc.density_ratio_predict(state64, g["test_probs"].astype(np.float64), g["test_prox"])
    np.testing.assert_allclose(cal64, g["f64_conf_cal"], rtol=1e-12, atol=0)


def test_isotonic_calibrators_match_reference(golden):
    """MultiIsotonicRegression / BinMeanShift outputs recorded from the reference classes (oracle/make_golden.py)."""
    g, d = golden("isotonic"), golden("density_ratio")
    vp, tp = d["val_probs"].astype(np.float64), d["test_probs"].astype(np.float64)
    rv, rt = g["rows_val"], g["rows_test"]
    val_out, cal = orc.multi_isotonic_fit_transform(vp, d["val_labels"])
    np.testing.assert_allclose(val_out[rv], g["val_out"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(orc.multi_isotonic_transform(cal, tp)[rt], g["test_out"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(cal.X_thresholds_, g["x_thresholds"], rtol=0, atol=0)
    for strategy in ("quantile", "uniform"):
        b_val, state = orc.bin_mean_shift_fit_transform(vp, d["val_prox"], d["val_labels"], 5, strategy)
        np.testing.assert_allclose(state[0], g[f"bms_{strategy}_edges"], rtol=1e-15)
        np.testing.assert_allclose(b_val[rv], g[f"bms_{strategy}_val_out"], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(orc.bin_mean_shift_transform(state, tp, d["test_prox"])[rt],
                                   g[f"bms_{strategy}_test_out"], rtol=1e-13, atol=1e-15)


def test_netcal_restatement_reproduces_its_fixture(golden):
    """The oracle's statement of netcal's one-vs-all HistogramBinning / IsotonicRegression (PARITY WITH NETCAL UNPINNED:
    the package is not installed) against the fixture make_golden.py wrote by running the reference's unmodified
    BinMeanShift around it; plus the properties the scheme guarantees: rows sum to 1, columns of unseen classes are 0,
    histogram-binning outputs only take values of the fitted bin maps."""
    g, d = golden("netcal_binning"), golden("density_ratio")
    vp, tp = d["val_probs"].astype(np.float64), d["test_probs"].astype(np.float64)
    for name, cls, kw in (("histogram_binning", orc.NetcalHistogramBinningCC, {"bins": 10}),
                          ("isotonic_regression", orc.NetcalIsotonicRegressionCC, {})):
        cal = cls(**kw).fit(vp, d["val_labels"])
        out = cal.transform(tp)
        np.testing.assert_allclose(out[g["rows_test"]], g[f"{name}_test_out"], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(out.sum(1), 1.0, rtol=1e-12)
        raw = cls(independent_probabilities=True, **kw).fit(vp, d["val_labels"]).transform(tp)
        assert raw.min() >= 0.0 and raw.max() <= 1.0
    y = d["val_labels"].copy()
    y[y == 3] = 4                                              # class 3 never occurs: no sub-model, column of zeros
    assert np.all(orc.NetcalHistogramBinningCC(bins=10).fit(vp, y).transform(tp)[:, 3] == 0.0)
    hb = orc.NetcalHistogramBinningCC(bins=4)
    got = hb.fit_transform(np.array([0.0, 0.1, 0.25, 0.3, 1.0]), np.array([0, 1, 1, 0, 1]))
    # bins [0, .25) -> mean(0, 1) = .5; [.25, .5) -> mean(1, 0) = .5; [.5, .75) empty -> centre .625; [.75, 1] -> 1
    np.testing.assert_allclose(hb.models[0][1], [0.5, 0.5, 0.625, 1.0])
    np.testing.assert_allclose(got, [0.5, 0.5, 0.5, 0.5, 1.0])


def test_oracle_float16_arithmetic_fit_matches_reference_golden(golden):
    """The oracle keeps the input dtype like the reference (float16 arrays stay float16 through norm / sum / exp):
    its class_confidence for float16 features equals the fixture the real reference produced (dac_float16.npz)."""
    from clip_calibration_b200 import synth
    g = golden("dac_float16")
    N, C, B, D, _, _ = synth.CONFIGS["sun397_l14"]
    txt_zs, txt_tuned, _ = synth.make_text(C, D, 0, rounding=synth.round_to_fp16)
    zs16, tu16 = txt_zs.astype(np.float16), txt_tuned.astype(np.float16)
    cc, *_ = orc.dac_fit(zs16[:B], zs16, tu16[:B], tu16, 5)
    assert np.array_equal(np.asarray(cc, np.float64), g["sun397_l14_cc16_k5"])


def test_oracle_fit_at_in21k_shape_matches_reference_golden(golden):
    from clip_calibration_b200 import synth
    g = golden("in21k_fit")
    txt_zs, txt_tuned, _ = synth.make_text(int(g["C"]), int(g["D"]), int(g["seed"]))
    B, sel = int(g["B"]), g["sel"][:24]
    cc, *_ = orc.dac_fit(txt_zs[:B], txt_zs[sel], txt_tuned[:B], txt_tuned[sel], int(g["k"]))
    assert np.array_equal(cc, g["cc"][:24])
