"""GPU (-m gpu): parity of the CUDA path, called through the C ABI, against the oracle and the
golden vectors produced by the real reference functions.

Tolerances (BASELINE.json north_star): predicted labels and kNN indices bit-exact except ties
(a "tie row" = reference top-2 logit gap <= 4e-5 * s/100: fp32 accumulation order alone can
flip those); confidences within 1e-4 relative; ECE within 1e-5 absolute; bin counts exact
except samples within 1e-6 of a bin edge.
"""
import os

import numpy as np
import pytest
import torch

from clip_calibration_b200 import native, pipeline, synth
from clip_calibration_b200 import table_math as tm
from clip_calibration_b200.tools import metrics
from clip_calibration_b200.trainers.calibration.distanse_aware_calibration import DistanseAwareCalibration
from clip_calibration_b200.trainers.calibration import tempscaling, proximity, vl_calibrator
from oracle import cpu_oracle as orc

pytestmark = pytest.mark.gpu

TIE_GAP = 4e-5


def dev(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    return t.to(dtype) if dtype is not None else t


def near_edge(conf, n_bins, tol=1e-6):
    edges = np.linspace(0, 1, n_bins + 1)
    return np.min(np.abs(conf[:, None].astype(np.float64) - edges[None, :]), axis=1) <= tol


# ----------------------------------------------------------------------------- K3
def test_bin_stats_edge_cases_exact(cuda_lib, golden):
    g = golden("metric_edge_cases")
    for name in g["names"]:
        conf, pred, gt = g[f"{name}_conf"], g[f"{name}_pred"], g[f"{name}_gt"]
        for nb in (10, 15):
            tab = metrics.bin_stats(conf, pred, gt, nb)
            assert np.array_equal(tab, orc.bin_table(conf, pred, gt, tm.uniform_thresholds(nb))), (name, nb)
            assert abs(metrics.ECE(conf, pred, gt, nb) - float(g[f"{name}_ece{nb}"])) < 1e-7, name
            assert abs(metrics.MCE(conf, pred, gt, nb) - float(g[f"{name}_mce{nb}"])) < 1e-7, name
            assert abs(metrics.AdaptiveECE(conf, pred, gt, nb) - float(g[f"{name}_ace{nb}"])) < 1e-7, (name, nb)
    assert metrics.ECE(np.array([1.0, 0.5], np.float32), np.array([0, 0]), np.array([0, 0]), 10) == 0.25
    assert isinstance(metrics.ECE(g["single_conf"], g["single_pred"], g["single_gt"]), np.float64)


def test_bin_stats_large_random_and_order_statistics(cuda_lib):
    rng = np.random.default_rng(11)
    n = 1_000_003
    conf = np.where(rng.random(n) < 0.05, 1.0, rng.random(n) ** 2).astype(np.float32)
    pred = rng.integers(0, 7, n).astype(np.int32)
    gt = rng.integers(0, 7, n)
    tab = metrics.bin_stats(conf, pred, gt, 15)
    assert np.array_equal(tab, orc.bin_table(conf, pred, gt, tm.uniform_thresholds(15)))
    ranks = [0, 1, 17, n // 3, n // 2, n - 2, n - 1]
    got = native.order_statistics(dev(conf), ranks)
    assert np.array_equal(got, np.sort(conf)[ranks])
    small = conf[:150000]
    assert abs(metrics.AdaptiveECE(small, pred[:150000], gt[:150000], 10)
               - orc.adaptive_ece(small, pred[:150000], gt[:150000], 10)) < 1e-7


def test_adaptive_ece_on_float64_plateaus(cuda_lib):
    """Calibrated confidences are float64 (isotonic / density-ratio outputs: a plateau value + 1e-9 * p): thousands of
    values that differ far below float32 resolution.  KBinsDiscretizer works on the float64 column - its edges fall
    INSIDE plateaus - so the order statistics and the binning must be float64 too."""
    rng = np.random.default_rng(17)
    n = 60_000
    plateau = rng.choice(np.array([0.12, 0.31, 0.48, 0.52, 0.77, 0.93]), size=n, p=[0.1, 0.15, 0.2, 0.25, 0.2, 0.1])
    conf = plateau + 1e-9 * rng.random(n)
    assert len(np.unique(conf.astype(np.float32))) <= 6 and len(np.unique(conf)) > n // 2
    pred = rng.integers(0, 5, n)
    gt = np.where(rng.random(n) < plateau, pred, (pred + 1) % 5)
    ranks = [0, 7, n // 10, n // 3, n // 2, n - 2, n - 1]
    got = native.order_statistics(torch.from_numpy(conf).cuda(), ranks)
    assert got.dtype == np.float64 and np.array_equal(got, np.sort(conf)[ranks])
    for nb in (10, 15):
        want = orc.adaptive_ece(conf, pred, gt, nb)
        assert abs(metrics.AdaptiveECE(conf, pred, gt, nb) - want) < 1e-9, nb
        # rounding the column to float32 first is a different (coarser) binning - the deviation the fix removes
    from clip_calibration_b200.evaluators import vl_evaluator
    probs = np.full((n, 5), 0.0)
    probs[np.arange(n), pred] = conf
    res = vl_evaluator.evaluate(probs, gt)
    assert abs(res["ace"] - 100.0 * orc.adaptive_ece(conf, pred, gt, 10)) < 1e-7


def test_piece_matches_reference(cuda_lib, golden):
    g = golden("proximity_piece")
    prox = np.exp(-np.mean(g["knn"], axis=-1))
    for nb in (10, 5):
        assert abs(metrics.PIECE(g["conf"], prox, g["pred"], g["labels"], nb, 10) - float(g[f"piece{nb}"])) < 1e-7
        assert abs(metrics.PIECE(g["conf"], prox, g["pred"], g["labels"], nb, 10, knn_strategy="uniform")
                   - float(g[f"piece{nb}_uniform"])) < 1e-7
    assert metrics.PIECE(g["conf"], np.full(len(prox), 0.5, np.float32), g["pred"], g["labels"], 10, 10, knn_strategy="uniform") \
        == metrics.ECE(g["conf"], g["pred"], g["labels"], 10)           # constant proximity: one proximity bin
    with pytest.raises(ValueError, match="kmeans"):
        metrics.PIECE(g["conf"], prox, g["pred"], g["labels"], 10, 10, knn_strategy="kmeans")


# ----------------------------------------------------------------------------- K1
@pytest.mark.parametrize("name,ks", [("eurosat", (5,)), ("sun397_l14", (1, 5, 10)), ("imagenet", (5,)), ("openvocab", (5,))])
def test_dac_fit_matches_reference(cuda_lib, name, ks, golden, synth_case):
    g, case = golden(name), synth_case(name)
    for k in ks:
        dac = DistanseAwareCalibration()
        dac.fit(case.base_zs, case.txt_zs, case.base_tuned, case.txt_tuned, k)
        cc_ref = g[f"cc_k{k}"]
        assert dac.class_confidence.dtype == np.float64 and dac.class_confidence.shape == cc_ref.shape
        np.testing.assert_allclose(dac.class_confidence, cc_ref, rtol=2e-6)
        assert np.array_equal(dac.class_confidence == 1.0, cc_ref == 1.0)        # base-class test agrees
        sel = g[f"fit_sel_k{k}"]
        kk = min(k, case.n_base)
        dref = g[f"knn_dist_tuned_k{k}"]
        dgot = dac.knn_distances_tuned.cpu().numpy()[sel][:, :kk]
        np.testing.assert_allclose(dgot, dref, rtol=2e-6, atol=1e-7)
        # indices: exact, except where two candidate distances coincide to fp32 rounding
        igot = dac.knn_indices_tuned.cpu().numpy()[sel][:, :kk]
        iref = g[f"knn_idx_tuned_k{k}"]
        diff = igot != iref
        if diff.any():
            assert np.all(np.abs(dgot[diff] - dref[diff]) <= 2e-6 * np.abs(dref[diff]) + 1e-7)
            assert diff.mean() < 1e-3
        zgot = dac.knn_indices_zs.cpu().numpy()[sel][:, :kk]
        assert (zgot != g[f"knn_idx_zs_k{k}"]).mean() < 1e-3


def test_dac_fit_at_the_in21k_shape(cuda_lib, golden):
    """BASELINE.json configs[4]'s own fit: 21,841 test classes x 10,000 base classes x 768-d (tensor-core filter +
    exact verification), against the reference's class_confidence on 256 classes spread over the vocabulary."""
    g = golden("in21k_fit")
    C, B, D, k = int(g["C"]), int(g["B"]), int(g["D"]), int(g["k"])
    txt_zs, txt_tuned, _ = synth.make_text(C, D, int(g["seed"]))
    assert abs(txt_tuned.astype(np.float64).sum() - float(g["txt_checksum"])) < 1e-6, "generator drifted"
    dac = DistanseAwareCalibration()
    dac.fit(txt_zs[:B], txt_zs, txt_tuned[:B], txt_tuned, k)
    sel = g["sel"]
    np.testing.assert_allclose(dac.class_confidence[sel], g["cc"], rtol=2e-6)
    assert np.array_equal(dac.class_confidence[sel] == 1.0, g["cc"] == 1.0)
    assert np.all(dac.class_confidence[:B] == 1.0)                     # every base class finds itself at distance 0
    dgot = dac.knn_distances_tuned.cpu().numpy()[sel]
    np.testing.assert_allclose(dgot, g["knn_dist_tuned"], rtol=2e-6, atol=1e-7)
    igot = dac.knn_indices_tuned.cpu().numpy()[sel]
    diff = igot != g["knn_idx_tuned"]
    if diff.any():                                                     # only where two distances coincide to rounding
        assert np.all(np.abs(dgot[diff] - g["knn_dist_tuned"][diff]) <= 2e-6 * g["knn_dist_tuned"][diff] + 1e-7)
    assert (dac.knn_indices_zs.cpu().numpy()[sel] != g["knn_idx_zs"]).mean() < 1e-3


@pytest.mark.parametrize("name,ks", [("sun397_l14", (1, 5, 10)), ("openvocab", (5,)), ("in21k", (5,))])
def test_dac_fit_in_the_reference_float16_arithmetic(cuda_lib, name, ks, golden):
    """The reference's default precision is fp16 and numpy keeps float16 through the whole fit
    (distanse_aware_calibration.py:28-42): `arithmetic="input"` must give the reference's class_confidence for
    float16 inputs exactly (half-precision values), where the default float32 arithmetic differs by ~1e-3."""
    g = golden("dac_float16")
    N, C, B, D, _, _ = synth.CONFIGS[name]
    txt_zs, txt_tuned, _ = synth.make_text(C, D, 0, rounding=synth.round_to_fp16)
    zs16, tu16 = txt_zs.astype(np.float16), txt_tuned.astype(np.float16)
    sel = g[f"{name}_sel"]
    for k in ks:
        want = g[f"{name}_cc16_k{k}"]
        dac = DistanseAwareCalibration()
        dac.fit(zs16[:B], zs16[sel], tu16[:B], tu16[sel], k, arithmetic="input")
        got = np.asarray(dac.class_confidence)
        assert got.dtype == np.float64
        off = got != want
        # np.exp's float32 result may sit within rounding of a half-way point of the float16 grid: allow a stray ulp
        assert off.mean() <= 0.005 and np.all(np.abs(got[off] - want[off]) <= 1.0e-3 * want[off]), (k, off.sum())
        kk = min(k, B)
        np.testing.assert_array_equal(dac.knn_distances_tuned.cpu().numpy()[:, :kk], g[f"{name}_dist16_tuned_k{k}"])
        # the default (float32) arithmetic is the accurate answer and is NOT what the reference gets from fp16 arrays
        dac32 = DistanseAwareCalibration()
        dac32.fit(zs16[:B], zs16[sel], tu16[:B], tu16[sel], k)
        rel = np.max(np.abs(np.asarray(dac32.class_confidence) - want) / want)
        assert 1e-5 < rel < 2e-3, rel
        assert abs(rel - float(g[f"{name}_rel_f32_vs_f16_k{k}"])) < 2e-5
    # torch.float16 CUDA tensors take the same route; float32 inputs ignore the switch
    dac = DistanseAwareCalibration()
    dac.fit(*[torch.from_numpy(x).cuda() for x in (zs16[:B], zs16[sel], tu16[:B], tu16[sel])], ks[-1], arithmetic="input")
    assert np.array_equal(np.asarray(dac.class_confidence), got)
    with pytest.raises(ValueError):
        dac.fit(zs16[:B], zs16[sel], tu16[:B], tu16[sel], 5, arithmetic="float16")


def test_dac_fit_k_larger_than_base_and_duplicates(cuda_lib):
    case = synth.make_case("tiny", 16, 6, 3, 64, 5, 0.3, seed=9)
    base_zs = case.base_zs.copy(); base_zs[2] = base_zs[1]            # duplicate base row (equal distances)
    dac = DistanseAwareCalibration()
    dac.fit(base_zs, case.txt_zs, case.base_tuned, case.txt_tuned, 5)   # k=5 > B=3
    cc, iz, it, dz, dt = orc.dac_fit(base_zs, case.txt_zs, case.base_tuned, case.txt_tuned, 5)
    np.testing.assert_allclose(dac.class_confidence, cc, rtol=2e-6)
    assert np.array_equal(dac.knn_indices_tuned.cpu().numpy()[:, :3], it)
    assert np.all(dac.knn_indices_tuned.cpu().numpy()[:, 3:] == -1)
    assert np.array_equal(dac.knn_indices_zs.cpu().numpy()[:, :3], iz)        # ties -> lowest index first


@pytest.mark.parametrize("nr,nq,d,k,drop", [(1000, 5000, 512, 5, False), (3000, 3000, 768, 10, True), (257, 4100, 64, 16, False),
                                             (10000, 2048, 512, 1, False)])
@pytest.mark.parametrize("stored_as", [None, torch.bfloat16, torch.float16], ids=["fp32", "bf16_values", "fp16_values"])
def test_knn_tensor_core_filter_equals_exhaustive_scan(cuda_lib, nr, nq, d, k, drop, stored_as):
    """ccal_knn_l2 (tcgen05 GEMM filter + exact verification + exhaustive redo of unproven rows) must
    return what the exhaustive fp32 scan returns - in all three operand modes of the filter: general fp32 values
    (bf16 hi/lo split, three MMAs), values that are exactly bf16 and values that are exactly fp16 (one MMA)."""
    g = torch.Generator(device="cuda").manual_seed(nr + nq)
    ref = torch.nn.functional.normalize(torch.randn(nr, d, device="cuda", generator=g) + 2.0, dim=-1)
    qry = None if drop else torch.nn.functional.normalize(torch.randn(nq, d, device="cuda", generator=g) + 2.0, dim=-1)
    if stored_as is not None:              # features cached in a 16-bit dtype (the reference's default is fp16)
        ref = ref.to(stored_as).float()
        qry = None if drop else qry.to(stored_as).float()
    qry = ref if drop else qry
    d_tc, i_tc = native.knn_l2(ref, qry, k, drop)
    d_ex, i_ex = native.knn_l2(ref, qry, k, drop, exhaustive=True)
    torch.testing.assert_close(d_tc, d_ex, rtol=2e-6, atol=2e-7)
    diff = i_tc != i_ex
    assert diff.float().mean() < 1e-3
    if diff.any():                                         # only where two distances coincide to rounding
        assert float((d_tc[diff] - d_ex[diff]).abs().max()) <= 1e-6
    # oracle spot check (reference formula on the CPU)
    rows = slice(0, 64)
    ref_d = orc.knn_dists(ref.cpu().numpy(), qry[rows].cpu().numpy(), k, drop_self=drop)
    np.testing.assert_allclose(d_tc[rows].cpu().numpy(), ref_d, rtol=3e-6, atol=3e-7)


@pytest.mark.parametrize("nr,nq,d,k,drop", [(5, 10, 512, 5, False), (199, 397, 768, 10, False), (500, 1000, 512, 5, False),
                                             (64, 16000, 64, 16, False), (1, 50, 4, 1, False), (300, 300, 128, 5, True),
                                             (3, 9, 1024, 7, False), (1000, 1000, 20, 16, True)])
def test_knn_small_problems_equal_exhaustive_scan(cuda_lib, nr, nq, d, k, drop):
    """Small shapes (below / around the switch between the exhaustive scan and the tensor-core filter): ccal_knn_l2 must
    return what the exhaustive scan returns, ties to the lowest index, k > nr padded."""
    g = torch.Generator(device="cuda").manual_seed(nr * 7 + nq)
    ref = torch.nn.functional.normalize(torch.randn(nr, d, device="cuda", generator=g) + 1.0, dim=-1)
    if nr > 4:
        ref[3] = ref[1]                                                # duplicate reference rows: equal distances
    qry = ref if drop else torch.nn.functional.normalize(torch.randn(nq, d, device="cuda", generator=g) + 1.0, dim=-1)
    d_s, i_s = native.knn_l2(ref, qry, k, drop)
    d_ex, i_ex = native.knn_l2(ref, qry, k, drop, exhaustive=True)
    torch.testing.assert_close(d_s, d_ex, rtol=2e-6, atol=2e-7)
    diff = i_s != i_ex
    if diff.any():                                         # only where two distances coincide to rounding
        assert float((d_s[diff] - d_ex[diff]).abs().max()) <= 1e-6
    want = orc.knn_dists(ref.cpu().numpy(), qry[:32].cpu().numpy(), min(k, nr - (1 if drop else 0)), drop_self=drop)
    np.testing.assert_allclose(d_s[:32, : want.shape[1]].cpu().numpy(), want, rtol=3e-6, atol=3e-7)


def test_knn_tensor_core_falls_back_on_ties_and_wild_norms(cuda_lib):
    g = torch.Generator(device="cuda").manual_seed(5)
    base = torch.nn.functional.normalize(torch.randn(40, 128, device="cuda", generator=g), dim=-1)
    ref = base.repeat_interleave(32, dim=0).contiguous()           # every row 32 times: ties across the cut
    qry = torch.nn.functional.normalize(torch.randn(2000, 128, device="cuda", generator=g), dim=-1)
    d_tc, i_tc = native.knn_l2(ref, qry, 5)
    d_ex, i_ex = native.knn_l2(ref, qry, 5, exhaustive=True)
    assert torch.equal(i_tc, i_ex)                                 # unproven rows are redone exhaustively (ties -> lowest index)
    torch.testing.assert_close(d_tc, d_ex, rtol=2e-6, atol=2e-7)
    # un-normalised rows with norms spread over 3 decades
    scale = torch.logspace(-1, 2, 1500, device="cuda")[:, None]
    ref2 = torch.randn(1500, 256, device="cuda", generator=g) * scale
    qry2 = torch.randn(1024, 256, device="cuda", generator=g) * 3.0
    d_tc, i_tc = native.knn_l2(ref2, qry2, 5)
    d_ex, i_ex = native.knn_l2(ref2, qry2, 5, exhaustive=True)
    torch.testing.assert_close(d_tc, d_ex, rtol=3e-6, atol=1e-6)
    assert (i_tc != i_ex).float().mean() < 1e-3


def test_proximity_knn(cuda_lib, golden):
    g = golden("proximity_piece")
    case = synth.make_case("prox", 600, 40, 20, 512, 5, 0.3, seed=3)
    val = synth.make_case("proxval", 300, 40, 20, 512, 5, 0.3, seed=4).img
    np.testing.assert_allclose(proximity.get_knn_dists(val, case.img, 5), g["knn"], rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(proximity.get_val_image_knn_dists(val, 5), g["knn_self"], rtol=2e-6, atol=2e-7)


# ----------------------------------------------------------------------------- K4
@pytest.mark.parametrize("n,c", [(1, 1), (77, 10), (300, 397), (129, 2048), (65, 2049), (33, 49408), (1000, 1000),
                                 (513, 256), (257, 512), (100, 132), (4099, 1028)])
def test_materialised_logits_dropins(cuda_lib, n, c):
    rng = np.random.default_rng(n * 1000 + c)
    logits = (rng.standard_normal((n, c)) * 8).astype(np.float32)
    if c > 3:
        logits[0, 3] = logits[0, 1] = logits[0].max() + 1.0               # a tie: first index wins
    cc = (0.9 + 0.1 * rng.random(c)).astype(np.float32)
    dac = DistanseAwareCalibration()
    dac.class_confidence = cc.astype(np.float64)
    before = logits.copy()
    out = dac.predict(logits.astype(np.float64))
    assert out.dtype == np.float32 and np.array_equal(logits, before)
    assert np.array_equal(out, orc.dac_predict(logits, cc))                # fp32 multiply: bit exact
    probs_ref = orc.softmax_lastaxis(orc.dac_predict(logits, cc))
    pref, cref = orc.pred_and_conf(probs_ref)
    pred, conf = dac.predict_confidence(logits)
    assert np.array_equal(pred, pref)
    np.testing.assert_allclose(conf, cref, rtol=1e-5)
    cal = vl_calibrator.VLCalibration(None, dac_flag=True)
    cal.dac_calibrator = dac
    probs = cal.predict(logits, np.zeros(n))
    np.testing.assert_allclose(probs, probs_ref, rtol=2e-5, atol=1e-30)


# ----------------------------------------------------------------------------- K2
def check_scoring(case, cc, pred_ref, conf_ref, gap_ref, ece_ref, counts_ref, dtype=torch.bfloat16):
    img, txt = dev(case.img, dtype), dev(case.txt_tuned, dtype)
    labels = dev(case.labels)
    ccd = dev(cc.astype(np.float32)) if cc is not None else None
    thr = tm.uniform_thresholds(10)
    table = native.new_table(10)
    pred, conf, rowmax = native.score_fused(img, txt, ccd, case.logit_scale, labels, thr, table, want_rowmax=True)
    torch.cuda.synchronize()
    pred, conf = pred.cpu().numpy(), conf.cpu().numpy()
    ties = gap_ref <= TIE_GAP * case.logit_scale / 100.0
    assert np.array_equal(pred[~ties], pred_ref[~ties]), f"{(pred[~ties] != pred_ref[~ties]).sum()} label mismatches"
    assert ties.mean() < 2e-3
    np.testing.assert_allclose(conf[~ties], conf_ref[~ties], rtol=1e-4)
    tab = native.table_to_numpy(table)
    assert np.array_equal(tab, orc.bin_table(conf, pred, case.labels, thr)), "fused binning != binning of its own outputs"
    assert abs(tm.ece_from_table(tab) - ece_ref) < 1e-5
    counts = tab[:, 0].astype(np.int64); folded = counts[:10].copy(); folded[9] += counts[10]
    moved = np.abs(folded - counts_ref).sum()
    assert moved <= 2 * (near_edge(conf_ref, 10).sum() + ties.sum()), (folded, counts_ref)
    return pred, conf


@pytest.fixture(params=["1", "2"], ids=["cta_group1", "cta_pair"])
def score_ctas(request, monkeypatch):
    """Run the fused kernel both as single CTAs (tcgen05 cta_group::1) and as CTA pairs (cta_group::2);
    by default the library picks pairs only for shards of >= 18,944 rows."""
    monkeypatch.setenv("CCAL_SCORE_CTAS", request.param)
    return request.param


@pytest.mark.parametrize("name", ["eurosat", "sun397_l14", "imagenet", "openvocab"])
def test_fused_scoring_matches_reference(cuda_lib, name, golden, synth_case, score_ctas):
    g, case = golden(name), synth_case(name)
    check_scoring(case, g["cc_k5"], g["dac_pred"], g["dac_conf"], g["dac_gap"], float(g["dac_ece10"]), g["dac_counts10"])
    check_scoring(case, None, g["nodac_pred"], g["nodac_conf"], g["nodac_gap"], float(g["nodac_ece10"]), g["nodac_counts10"])


@pytest.mark.parametrize("n,c,d", [(1, 1, 64), (127, 255, 64), (129, 257, 128), (1000, 513, 512), (300, 40, 1024), (4096, 3000, 640)])
def test_fused_scoring_ragged_shapes_fp16_and_bf16(cuda_lib, n, c, d, score_ctas):
    for dtype, rounding in ((torch.bfloat16, synth.round_to_bf16), (torch.float16, synth.round_to_fp16)):
        case = synth.make_case("ragged", n, c, max(1, c // 2), d, 5, 0.3, seed=n + c, rounding=rounding)
        cc = (0.95 + 0.05 * np.random.default_rng(c).random(c)).astype(np.float32)
        pref, cref, gap = orc.score_chain(case.img, case.txt_tuned, cc, case.logit_scale)
        ece_ref = orc.ece(cref, pref, case.labels, 10)
        counts = np.histogram(cref, np.linspace(0, 1, 11))[0]
        check_scoring(case, cc, pref, cref, gap, ece_ref, counts, dtype)


def _device_case(n, c, d, signal, dtype, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    unit = lambda x: x / x.norm(dim=-1, keepdim=True)
    u = unit(torch.randn(d, device="cuda", generator=g))
    txt = unit(u[None, :] + 0.6 / d ** 0.5 * torch.randn(c, d, device="cuda", generator=g)).to(dtype)
    labels = torch.randint(0, c, (n,), device="cuda", generator=g)
    img = torch.empty((n, d), dtype=dtype, device="cuda")
    for lo in range(0, n, 65536):
        hi = min(n, lo + 65536)
        raw = signal * txt[labels[lo:hi]].float() + torch.randn(hi - lo, d, device="cuda", generator=g) / d ** 0.5
        img[lo:hi] = unit(raw).to(dtype)
    cc = (0.97 + 0.03 * torch.rand(c, device="cuda", generator=g)).float()
    cc[: c // 3] = 1.0
    return img, txt.contiguous(), labels, cc


@pytest.mark.parametrize("n,c,d,signal,dtype,use_cc", [
    (3000, 2048, 512, 0.4, torch.bfloat16, True),        # few rows: the redo kernel cuts its tiles into class ranges
    (50_000, 8192, 512, 0.4, torch.bfloat16, True),
    (60_000, 21841, 768, 0.45, torch.bfloat16, True),    # d = 768: streaming verify pass, 2-stage redo kernel
    (33_333, 5000, 640, 0.3, torch.bfloat16, True),      # ragged everything
    (30_000, 4096, 256, 0.3, torch.float16, True),
    (20_000, 3000, 128, 0.3, torch.bfloat16, False),     # no multipliers: redo only where the guess missed the maximum
    (320_000, 2048, 128, 0.02, torch.bfloat16, True),    # near-random labels: > 32,768 rows redone (no class ranges)
])
def test_fp8_guess_pipeline_is_bit_identical_to_two_pass(cuda_lib, n, c, d, signal, dtype, use_cc, monkeypatch):
    """ccal_score_fused through the FP8-guess -> exact-logit -> bf16-verify -> redo pipeline (what large shards run)
    against the plain two-pass kernel on the same inputs: labels, confidences, row maxima and the bin table are
    identical bit for bit, whatever the guess quality; a row's result does not depend on how the shard is cut."""
    img, txt, labels, cc = _device_case(n, c, d, signal, dtype, seed=n + c)
    if not use_cc:
        cc = None
    thr = tm.uniform_thresholds(15)
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("CCAL_SCORE_FP8", mode)
        native.score_guess_stats(reset=True)
        table = native.new_table(15)
        pred, conf, rowmax = native.score_fused(img, txt, cc, 100.0, labels, thr, table, want_rowmax=True)
        outs[mode] = (pred, conf, rowmax, table, native.score_guess_stats(reset=True))
    p0, c0, r0, t0, s0 = outs["0"]
    p1, c1, r1, t1, s1 = outs["1"]
    assert s0 == (0, 0) and s1[0] == n
    assert 0 < s1[1] < 0.6 * n, f"redo count {s1[1]} of {n}: the FP8 guess or the exact guessed-class logit is off"
    assert torch.equal(p0, p1) and torch.equal(c0, c1) and torch.equal(r0, r1) and torch.equal(t0, t1)
    assert tm.total_count(native.table_to_numpy(t1)) == n
    # outputs are optional, sub-shards give the same rows, and the table accumulates
    m = n // 3
    t2 = native.new_table(15)
    p2, c2, _ = native.score_fused(img[:m].contiguous(), txt, cc, 100.0, labels[:m].contiguous(), thr, t2)
    native.score_fused(img[m:].contiguous(), txt, cc, 100.0, labels[m:].contiguous(), thr, t2, want_pred=False, want_conf=False)
    assert torch.equal(p2, p1[:m]) and torch.equal(c2, c1[:m]) and torch.equal(t2, t1)
    monkeypatch.delenv("CCAL_SCORE_FP8")


def test_fp8_guess_pipeline_on_the_open_vocabulary_golden(cuda_lib, golden, synth_case, monkeypatch):
    """The reference-generated open-vocabulary fixture (1,024 rows x 49,408 classes) through the forced pipeline."""
    monkeypatch.setenv("CCAL_SCORE_FP8", "1")
    g, case = golden("openvocab"), synth_case("openvocab")
    native.score_guess_stats(reset=True)
    check_scoring(case, g["cc_k5"], g["dac_pred"], g["dac_conf"], g["dac_gap"], float(g["dac_ece10"]), g["dac_counts10"])
    check_scoring(case, None, g["nodac_pred"], g["nodac_conf"], g["nodac_gap"], float(g["nodac_ece10"]), g["nodac_counts10"])
    assert native.score_guess_stats(reset=True)[0] > 0
    monkeypatch.delenv("CCAL_SCORE_FP8")


@pytest.mark.parametrize("n,c,d,dtype", [(100, 49408, 512, torch.bfloat16), (1, 3000, 64, torch.float16), (1000, 21841, 768, torch.bfloat16),
                                         (4097, 5000, 512, torch.bfloat16), (300, 2049, 128, torch.float32)])
def test_column_split_mode_for_small_batches(cuda_lib, n, c, d, dtype, monkeypatch):
    """Few image rows x a large vocabulary (the reference's 100-image batches, serving-style queries): every
    row tile is cut into class ranges so that all SMs work on it.  Same labels as the unsplit kernel and the
    oracle, confidences to rounding, identical fused bin table semantics."""
    rounding = {torch.bfloat16: synth.round_to_bf16, torch.float16: synth.round_to_fp16,
                torch.float32: lambda x: np.asarray(x, np.float32)}[dtype]
    case = synth.make_case("small", n, c, max(1, c // 2), d, 5, 0.3, seed=n + c, rounding=rounding)
    cc = (0.95 + 0.05 * np.random.default_rng(c).random(c)).astype(np.float32)
    img, txt, ccd, lab = dev(case.img, dtype), dev(case.txt_tuned, dtype), dev(cc), dev(case.labels)
    thr = tm.uniform_thresholds(10)
    t_split, t_plain = native.new_table(10), native.new_table(10)
    p1, c1, r1 = native.score_fused(img, txt, ccd, 100.0, lab, thr, t_split, want_rowmax=True)
    monkeypatch.setenv("CCAL_SCORE_NOSPLIT", "1")
    p0, c0, r0 = native.score_fused(img, txt, ccd, 100.0, lab, thr, t_plain, want_rowmax=True)
    monkeypatch.delenv("CCAL_SCORE_NOSPLIT")
    assert torch.equal(p1, p0) and torch.equal(r1, r0) and torch.equal(c1, c0)      # canonical summation order
    pref, cref, gap = orc.score_chain(case.img, case.txt_tuned, cc, 100.0)
    ok = gap > (2e-4 if dtype == torch.float32 else TIE_GAP)
    assert np.array_equal(p1.cpu().numpy()[ok], pref[ok])
    np.testing.assert_allclose(c1.cpu().numpy()[ok], cref[ok], rtol=1e-4)
    tab = native.table_to_numpy(t_split)
    assert np.array_equal(tab, orc.bin_table(c1.cpu().numpy(), p1.cpu().numpy(), case.labels, thr))


def test_end_to_end_pipeline_and_host_path(cuda_lib, golden, synth_case):
    g, case = golden("imagenet"), synth_case("imagenet")
    scorer = pipeline.CalibratedScorer.from_dac(case.base_zs, case.txt_zs, case.base_tuned, case.txt_tuned, k=5,
                                                logit_scale=case.logit_scale, n_bins=10, operand_dtype=torch.bfloat16)
    pred, conf = scorer.score(case.img, case.labels)
    s = scorer.summary()
    assert s["n"] == len(case.labels)
    assert abs(s["ece"] - float(g["dac_ece10"])) < 1e-5 and abs(s["mce"] - float(g["dac_mce10"])) < 1e-5
    assert abs(s["accuracy"] - float(g["dac_acc"])) < 1e-4 and abs(s["confidence"] - float(g["dac_mean_conf"])) < 1e-5
    assert abs(metrics.AdaptiveECE(conf, pred, dev(case.labels), 10) - float(g["dac_ace10"])) < 1e-5
    # host inputs, chunked H2D overlapped with compute: identical table
    device_table = s["table"].copy()
    scorer.reset()
    host = torch.from_numpy(case.img).to(torch.bfloat16).pin_memory()
    scorer.accumulate_host(host, torch.from_numpy(case.labels).pin_memory(), chunk_rows=8192)
    assert np.array_equal(scorer.reduced_table(), device_table)
    # the per-batch loop of the reference (100 images at a time), buffered on the device
    scorer.reset()
    for lo in range(0, len(case.labels), 100):
        scorer.add(case.img[lo:lo + 100], case.labels[lo:lo + 100], flush_rows=16384)
    assert np.array_equal(scorer.reduced_table(), device_table)
    # shards add up exactly (what the multi-GPU all-reduce relies on)
    scorer.reset()
    for r in range(3):
        lo, hi = pipeline.shard_bounds(len(case.labels), r, 3)
        scorer.score(case.img[lo:hi], case.labels[lo:hi])
    assert np.array_equal(scorer.reduced_table(), device_table)


# ----------------------------------------------------------------------------- K5
@pytest.mark.parametrize("n,c,d,t", [(200, 37, 128, 4.6052), (1000, 500, 512, 4.0), (64, 1000, 768, 5.0)])
def test_temperature_scaling_loss_and_gradient(cuda_lib, n, c, d, t, score_ctas):
    case = synth.make_case("ts", n, c, max(1, c // 2), d, 5, 0.3, seed=n)
    loss, grad = tempscaling.ts_loss_and_grad(case.img, case.txt_tuned, case.labels, t)
    lref, gref = orc.ts_loss_and_grad(case.img, case.txt_tuned, case.labels, t)
    assert abs(loss - lref) <= 1e-4 * max(1.0, abs(lref)), (loss, lref)
    assert abs(grad - gref) <= 1e-4 * max(1.0, abs(gref)), (grad, gref)


def test_fit_logit_scale_reduces_loss_and_checkpoint_roundtrip(cuda_lib, tmp_path):
    case = synth.make_case("tsfit", 512, 100, 50, 512, 5, 0.3, seed=21)
    l0, _ = orc.ts_loss_and_grad(case.img, case.txt_tuned, case.labels, tempscaling.INIT_LOG_SCALE)
    t = tempscaling.fit_logit_scale(case.img, case.txt_tuned, case.labels, epochs=5)
    l1, g1 = orc.ts_loss_and_grad(case.img, case.txt_tuned, case.labels, t)
    assert l1 < l0
    # the objective's own minimiser: gradient ~ 0 there and no SGD trajectory beats it
    t_star = tempscaling.solve_logit_scale(case.img, case.txt_tuned, case.labels)
    l_star, g_star = orc.ts_loss_and_grad(case.img, case.txt_tuned, case.labels, t_star)
    assert abs(g_star) < 1e-4 and l_star <= l1 + 1e-9
    for dt in (-0.05, 0.05):
        assert orc.ts_loss_and_grad(case.img, case.txt_tuned, case.labels, t_star + dt)[0] > l_star
    tempscaling.save_logit_scale(str(tmp_path), t, 5)
    assert abs(tempscaling.load_logit_scale(str(tmp_path), 5) - t) < 1e-6
    learner = tempscaling.ScaleLearner(None, torch.float32)
    assert abs(float(learner()) - np.exp(4.6052)) < 1e-3 and list(learner.state_dict()) == ["logit_scale"]


def test_fit_logit_scale_device_loop_matches_host_stepping(cuda_lib):
    """f-3: with 16-bit operands the whole SGD schedule runs on the device (the scalar, its momentum buffer and the
    running loss live in device memory; no read-back until the end).  It must walk the same trajectory as stepping
    the same objective from the host, batch by batch, with the same shuffles."""
    import math
    case = synth.make_case("tsdev", 500, 120, 60, 512, 5, 0.3, seed=5)
    img = torch.from_numpy(case.img).cuda().to(torch.bfloat16)
    txt = torch.from_numpy(case.txt_tuned).cuda().to(torch.bfloat16)
    y = torch.from_numpy(case.labels).cuda()
    epochs, bs, lr, mom, wd = 4, 32, 0.05, 0.9, 5e-4
    n0 = native.launch_count()
    t_dev, hist = tempscaling.fit_logit_scale(img, txt, y, epochs=epochs, lr=lr, batch_size=bs, momentum=mom, weight_decay=wd,
                                              return_history=True)
    batches = epochs * (500 // bs)
    assert native.launch_count() - n0 == 3 * batches            # scoring kernel + fixed-order reduce + SGD step, per batch
    gen = torch.Generator().manual_seed(0)
    t, vel = tempscaling.INIT_LOG_SCALE, 0.0
    for epoch in range(epochs):
        cur = 1e-5 if epoch < 1 else 0.5 * lr * (1.0 + math.cos(math.pi * (epoch - 1) / (epochs - 1)))
        order = torch.randperm(500, generator=gen).cuda()
        for lo in range(0, 500 - bs + 1, bs):
            sel = order[lo:lo + bs]
            _, g = tempscaling.ts_loss_and_grad(img[sel], txt, y[sel], t)
            vel = mom * vel + g + wd * t
            t -= cur * vel
        assert abs(hist[epoch]["t"] - t) < 5e-5, (epoch, hist[epoch]["t"], t)
    assert abs(t_dev - t) < 5e-5 and len(hist) == epochs
    l0, _ = orc.ts_loss_and_grad(case.img, case.txt_tuned, case.labels, tempscaling.INIT_LOG_SCALE)
    l1, _ = orc.ts_loss_and_grad(case.img, case.txt_tuned, case.labels, t_dev)
    assert l1 < l0


def test_c_abi_from_plain_c(cuda_lib):
    """examples/c_abi_smoke.c: the library driven from C (cudaMalloc'ed buffers, its own stream, no torch)."""
    import subprocess
    from clip_calibration_b200 import build as _build
    exe = _build.C_DEMO_BIN
    assert os.path.exists(exe), "run python -m clip_calibration_b200.build (builds the C demo too)"
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "C ABI smoke: OK" in res.stdout


def test_nearest_tokens_matches_cdist_argsort(cuda_lib):
    """interpret_prompts/interpret_prompt.py:69-72 on the CUDA kNN: 16 context vectors vs a 49,408-token table."""
    from clip_calibration_b200 import interpret
    g = torch.Generator().manual_seed(0)
    emb = torch.randn(49408, 512, generator=g) * 0.02
    ctx = emb[torch.randint(0, 49408, (16,), generator=g)] + 0.01 * torch.randn(16, 512, generator=g)
    idx, dist = interpret.nearest_tokens(ctx.numpy(), emb.numpy(), 5)
    full = torch.cdist(ctx, emb)
    ref_idx = torch.argsort(full, dim=1)[:, :5]
    assert np.array_equal(idx, ref_idx.numpy())
    np.testing.assert_allclose(dist, torch.gather(full, 1, ref_idx).numpy(), rtol=2e-5, atol=1e-6)


def test_empty_inputs(cuda_lib):
    e_img = torch.zeros((0, 64), dtype=torch.bfloat16, device="cuda")
    txt = torch.zeros((5, 64), dtype=torch.bfloat16, device="cuda")
    pred, conf, _ = native.score_fused(e_img, txt)
    assert pred.numel() == 0 and conf.numel() == 0
    tab = native.bin_stats(torch.zeros(0, device="cuda"), torch.zeros(0, dtype=torch.int32, device="cuda"),
                           torch.zeros(0, dtype=torch.int64, device="cuda"), tm.uniform_thresholds(10))
    assert int(tab.sum()) == 0
    d, i = native.knn_l2(torch.randn(7, 64, device="cuda"), torch.zeros((0, 64), device="cuda"), 3)
    assert d.shape == (0, 3) and i.shape == (0, 3)
    p, c = native.logits_confidence(torch.zeros((0, 9), device="cuda"))
    assert p.numel() == 0 and c.numel() == 0


@pytest.mark.parametrize("n,c,d", [(2000, 300, 512), (3000, 1000, 768), (1024, 49408, 512), (20000, 500, 128)])
def test_fp32_features_split_precision_mode(cuda_lib, n, c, d, score_ctas):
    """fp32 features that are NOT representable in 16 bits: the default (split-precision: fp16 hi/lo pairs,
    3 MMAs per K step) must match the reference's fp32 path like the 16-bit path matches it on pre-rounded
    inputs - labels exact outside the tie band, confidences within 1e-4 relative - where plainly rounding the
    features to fp16 / bf16 does not."""
    ident = lambda x: np.asarray(x, np.float32)
    case = synth.make_case("fp32", n, c, max(1, c // 2), d, 5, 0.3, seed=n, rounding=ident)
    cc = (0.95 + 0.05 * np.random.default_rng(1).random(c)).astype(np.float32)
    pref, cref, gap = orc.score_chain(case.img, case.txt_tuned, cc, 100.0)
    dac = DistanseAwareCalibration()
    dac.class_confidence = cc.astype(np.float64)
    pred, conf = dac.predict_from_features(case.img, case.txt_tuned, 100.0)            # default -> split precision
    ok = gap > 2e-4
    assert np.array_equal(pred[ok], pref[ok]) and (~ok).mean() < 5e-3
    np.testing.assert_allclose(conf[ok], cref[ok], rtol=1e-4)
    # fused binning + TS objective go through the same operands
    table = native.new_table(10)
    native.score_fused(dev(case.img), dev(case.txt_tuned), dev(cc), 100.0, dev(case.labels), tm.uniform_thresholds(10), table)
    assert abs(tm.ece_from_table(native.table_to_numpy(table)) - orc.ece(cref, pref, case.labels, 10)) < 1e-5
    if n <= 3000:
        loss, grad = tempscaling.ts_loss_and_grad(case.img, case.txt_tuned, case.labels, 4.6052)
        lref, gref = orc.ts_loss_and_grad(case.img, case.txt_tuned, case.labels, 4.6052)
        assert abs(loss - lref) <= 2e-5 * max(1.0, abs(lref)) and abs(grad - gref) <= 2e-5 * max(1.0, abs(gref))
    # rounding the features instead is visibly worse (that is why it is not the default)
    p16, c16 = dac.predict_from_features(case.img, case.txt_tuned, 100.0, operand_dtype=torch.bfloat16)
    assert np.median(np.abs(c16[ok] - cref[ok]) / cref[ok]) > 20 * np.median(np.abs(conf[ok] - cref[ok]) / cref[ok])


# ----------------------------------------------------------------------------- errors
def test_bad_arguments_fail_loudly(cuda_lib):
    with pytest.raises(ValueError):
        native.score_fused(torch.zeros(8, 100, dtype=torch.bfloat16, device="cuda"),
                           torch.zeros(8, 100, dtype=torch.bfloat16, device="cuda"))         # D not a multiple of 64
    with pytest.raises(ValueError):
        native.score_fused(torch.zeros(8, 64, dtype=torch.float64, device="cuda"),
                           torch.zeros(8, 64, dtype=torch.float64, device="cuda"))           # fp64 operands
    with pytest.raises(ValueError):
        native.score_fused(torch.zeros(8, 64, dtype=torch.float16, device="cuda"),
                           torch.zeros(8, 64, dtype=torch.bfloat16, device="cuda"))          # mixed operand dtypes
    with pytest.raises(ValueError):
        native.knn_l2(torch.zeros(4, 64, device="cuda"), torch.zeros(4, 64, device="cuda"), 99)
    bad = DistanseAwareCalibration()
    bad.class_confidence = np.array([1.0, -0.5, 1.0])
    with pytest.raises(ValueError):
        bad.predict(np.zeros((2, 3), np.float32))
    with pytest.raises(Exception):
        native.bin_stats(torch.zeros(4), torch.zeros(4, dtype=torch.int32), torch.zeros(4, dtype=torch.int64), [0.5])


# ----------------------------------------------------------------------------- f-4 density-ratio calibration
def test_density_ratio_calibration_matches_reference_fixture(cuda_lib, golden):
    """DensityRatioCalibration.fit/.predict against what the reference class returned (tests/golden/density_ratio.npz,
    oracle/make_golden.py) and against the oracle, for float32 and float64 probability matrices."""
    from clip_calibration_b200.trainers.calibration.density_ratio_calibration import DensityRatioCalibration
    g = golden("density_ratio")
    cal = DensityRatioCalibration()
    cal.fit(g["val_probs"], g["val_preds"], g["val_labels"], g["val_prox"])
    np.testing.assert_allclose(cal.dens_true.bw, g["f32_bw_true"], rtol=1e-12)
    np.testing.assert_allclose(cal.dens_false.bw, g["f32_bw_false"], rtol=1e-12)
    assert float(cal.false_true_ratio) == float(g["f32_ratio"])
    before = g["test_probs"].copy()
    out = cal.predict(g["test_probs"], g["test_prox"])
    assert out.dtype == np.float64 and out.shape == before.shape and np.array_equal(g["test_probs"], before)
    # tolerance: ex2.approx (2e-7) on float64 exponents; the reference sums the other classes in float32 (1e-7)
    np.testing.assert_allclose(out, g["f32_probs_out"], rtol=2e-6, atol=1e-300)
    np.testing.assert_allclose(out.sum(axis=1), 1.0, rtol=1e-6)
    conf = cal.calibrated_confidence(g["test_probs"].max(axis=1), g["test_prox"])
    np.testing.assert_allclose(conf, g["f32_conf_cal"], rtol=2e-6, atol=1e-300)
    # float64 probabilities
    cal64 = DensityRatioCalibration()
    cal64.fit(g["val_probs"].astype(np.float64), g["val_preds"], g["val_labels"], g["val_prox"])
    out64 = cal64.predict(g["test_probs"].astype(np.float64), g["test_prox"])
    state = orc.density_ratio_fit(g["val_probs"].astype(np.float64), g["val_preds"], g["val_labels"], g["val_prox"])
    want64, _ = orc.density_ratio_predict(state, g["test_probs"].astype(np.float64), g["test_prox"])
    np.testing.assert_allclose(out64, want64, rtol=2e-6, atol=1e-300)


@pytest.mark.parametrize("m,n,bw", [(1, 5, (0.05, 0.002)), (7, 1, (0.1, 0.1)), (513, 700, (0.03, 0.0007)),
                                    (20000, 3000, (0.02, 0.001)), (300, 100000, (0.05, 0.003))])
def test_kde2_pdf_against_float64_oracle(cuda_lib, m, n, bw):
    """ccal_kde2_pdf against the float64 restatement, including far tails (densities down to 1e-300: the ratio of two
    such values is what the calibrator uses) and query points that coincide with data points."""
    rng = np.random.default_rng(m + n)
    data = np.stack([rng.random(m), 0.4 + 0.01 * rng.standard_normal(m)], axis=1)
    q = np.stack([rng.random(n), 0.4 + 0.02 * rng.standard_normal(n)], axis=1)
    q[: min(n, m, 3)] = data[: min(n, m, 3)]
    if n > 4:
        q[4] = (0.5, 0.4 + 37 * bw[1])                    # deep tail: exp(-684) ~ 1e-297 still representable
    dens = orc.KDEMultivariateCC([data[:, 0], data[:, 1]])
    dens.bw = np.array(bw)
    sel = np.unique(np.concatenate([np.arange(min(n, 64)), rng.integers(0, n, 64)]))
    want = dens.pdf(q[sel])
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    got = native.kde2_pdf(dev(data[:, 0]), dev(data[:, 1]), dev(q[:, 0]), dev(q[:, 1]), *bw).cpu().numpy()
    assert got.shape == (n,) and np.all(np.isfinite(got))
    np.testing.assert_allclose(got[sel], want, rtol=1e-6, atol=1e-305)


@pytest.mark.parametrize("n,c,dtype", [(1, 2, np.float32), (300, 10, np.float32), (77, 397, np.float64), (40, 2049, np.float32),
                                       (9, 49408, np.float32), (5000, 1000, np.float32)])
def test_density_ratio_apply_rows(cuda_lib, n, c, dtype):
    rng = np.random.default_rng(n + c)
    logits = rng.standard_normal((n, c)) * 3
    probs = orc.softmax_lastaxis(logits).astype(dtype)
    if c > 3:
        probs[0, 1] = probs[0, 3] = probs[0].max()                    # tie: the first maximum is the prediction
    t = rng.random(n); f = rng.random(n); ratio = 0.37
    if n > 2:
        t[1], f[1] = 0.0, 0.0                                          # both densities underflowed: eps floor -> 0
        t[2], f[2] = 1e-290, 3e-290
    dens_t = type("D", (), {"pdf": staticmethod(lambda d: t)})
    dens_f = type("D", (), {"pdf": staticmethod(lambda d: f)})
    want, wcal = orc.density_ratio_predict((dens_t, dens_f, ratio), probs, np.zeros(n))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    out, cal, pred = native.density_ratio_apply(dev(probs), dev(t), dev(f), ratio)
    assert np.array_equal(pred.cpu().numpy(), np.argmax(probs, axis=1))
    np.testing.assert_allclose(cal.cpu().numpy(), wcal, rtol=1e-14, atol=0)
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=1e-6 if dtype == np.float32 else 1e-12, atol=0)


def test_vl_calibration_scaling_based_with_proximity(cuda_lib, golden):
    """VLCalibration(base_calibration_mode='scaling_based', procal_flag=True): softmax -> density-ratio calibrator
    (reference vl_calibrator.py:95-96, :116-119)."""
    g = golden("density_ratio")
    val_logits = np.log(g["val_probs"].astype(np.float64))             # softmax(log p) == p
    test_logits = np.log(g["test_probs"].astype(np.float64))
    val_dict = {"val_logits": val_logits, "val_labels": g["val_labels"], "val_image_features": None,
                "val_text_features": None, "val_image_knn_dists": -np.log(g["val_prox"].astype(np.float64))[:, None]}
    cal = vl_calibrator.VLCalibration(None, base_calibration_mode="scaling_based", procal_flag=True, val_dict=val_dict)
    cal.fit()
    out = cal.predict(test_logits, g["test_prox"])
    assert out.dtype == np.float64
    np.testing.assert_allclose(out, g["f32_probs_out"], rtol=5e-5, atol=1e-12)   # probabilities re-derived through log/softmax in fp32
    nameless = vl_calibrator.VLCalibration(None, base_calibration_mode="bin_based", val_dict=val_dict)
    nameless.fit()                               # no calibrator name: the reference's if/elif chain builds nothing
    assert nameless.base_calibrator is None


# ----------------------------------------------------------------------------- macro-F1 (evaluator)
@pytest.mark.parametrize("n,c", [(1, 1), (1000, 10), (5000, 397), (200000, 2048), (300000, 49408), (0, 5)])
def test_class_counts_and_macro_f1(cuda_lib, n, c):
    import warnings
    from sklearn.metrics import f1_score
    rng = np.random.default_rng(n + c)
    gt = rng.integers(0, c, n)
    pred = np.where(rng.random(n) < 0.7, gt, rng.integers(0, c, n))
    counts = metrics.class_counts(pred.astype(np.int32), gt, n_classes=c)
    want = np.zeros((c, 3), np.int64)
    np.add.at(want[:, 0], gt[pred == gt], 1)
    np.add.at(want[:, 1], pred[pred != gt], 1)
    np.add.at(want[:, 2], gt[pred != gt], 1)
    assert np.array_equal(counts, want)
    if n:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = f1_score(gt, pred, average="macro", labels=np.unique(gt))      # evaluators/vl_evaluator.py:74-79
        assert abs(metrics.macro_f1(pred, gt) - ref) < 1e-12                     # int64 preds, class count inferred
        # accumulation over shards = one pass (what the multi-GPU path all-reduces)
        acc = native.class_counts(torch.from_numpy(pred[: n // 2]).cuda(), torch.from_numpy(gt[: n // 2]).cuda(), c)
        native.class_counts(torch.from_numpy(pred[n // 2:]).cuda(), torch.from_numpy(gt[n // 2:]).cuda(), c, acc)
        assert np.array_equal(acc.cpu().numpy(), want)


def test_evaluator_reports_reference_keys(cuda_lib, golden):
    from clip_calibration_b200.evaluators import vl_evaluator
    import warnings
    from sklearn.metrics import f1_score
    g = golden("eurosat")
    case = synth.make_config("eurosat", seed=0)
    res = vl_evaluator.evaluate_pred_conf(g["dac_pred"], g["dac_conf"], case.labels, 10)
    for key in ("accuracy", "error_rate", "macro_f1", "confidence", "ece", "mce", "ace"):
        assert key in res
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        f1 = 100.0 * f1_score(case.labels, g["dac_pred"], average="macro", labels=np.unique(case.labels))
    assert abs(res["macro_f1"] - f1) < 1e-9
    assert abs(res["ece"] - 100.0 * float(g["dac_ece10"])) < 1e-5


def test_scorer_evaluate_reports_every_reference_key(cuda_lib, golden):
    """CalibratedScorer(keep_outputs=True).evaluate(): the fused path must reproduce the evaluator's result dict
    (evaluators/vl_evaluator.py:59-116) computed from the reference fixture's per-image outputs."""
    import warnings
    from sklearn.metrics import f1_score
    from clip_calibration_b200.evaluators import vl_evaluator
    g = golden("sun397_l14")
    case = synth.make_config("sun397_l14", seed=0)
    prox = np.random.default_rng(3).random(case.img.shape[0]).astype(np.float32)
    scorer = pipeline.CalibratedScorer.from_dac(case.base_zs, case.txt_zs, case.base_tuned, case.txt_tuned, k=5,
                                                logit_scale=case.logit_scale, operand_dtype=torch.bfloat16,
                                                keep_outputs=True, group=False)
    n = case.img.shape[0]
    # three entry points, in order: per-batch add(), a device shard, a pinned host shard
    a, b = n // 3, 2 * n // 3
    for lo in range(0, a, 1000):
        scorer.add(case.img[lo:min(a, lo + 1000)], case.labels[lo:min(a, lo + 1000)], flush_rows=4096)
    scorer.flush()
    scorer.score(case.img[a:b], case.labels[a:b])
    scorer.accumulate_host(torch.from_numpy(case.img[b:]).to(torch.bfloat16).pin_memory(),
                           torch.from_numpy(case.labels[b:]).pin_memory(), chunk_rows=2048)
    res = scorer.evaluate(proximity=prox)
    want = vl_evaluator.evaluate_pred_conf(g["dac_pred"], g["dac_conf"], case.labels, 10, 10, proximity=prox)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        f1 = 100.0 * f1_score(case.labels, g["dac_pred"], average="macro", labels=np.unique(case.labels))
    assert abs(res["macro_f1"] - f1) < 1e-9
    for key in ("accuracy", "error_rate", "macro_f1", "confidence", "ece", "mce", "ace", "piece"):
        assert abs(res[key] - want[key]) < 2e-3, (key, res[key], want[key])      # percent; conf rel 1e-4 -> ECE 1e-5
    assert abs(res["ece"] - 100.0 * float(g["dac_ece10"])) < 1e-3
    assert abs(res["ace"] - 100.0 * float(g["dac_ace10"])) < 1e-3
    with pytest.raises(RuntimeError):
        pipeline.CalibratedScorer(case.txt_tuned, None, group=False).evaluate()


# ----------------------------------------------------------------------------- f-4 isotonic calibrators
def test_multi_isotonic_regression_matches_reference_fixture(cuda_lib, golden):
    from clip_calibration_b200.trainers.calibration.multi_isotonic_regression import MultiIsotonicRegression
    g, d = golden("isotonic"), golden("density_ratio")
    vp, tp = d["val_probs"].astype(np.float64), d["test_probs"].astype(np.float64)
    cal = MultiIsotonicRegression()
    val_out = cal.fit_transform(vp, d["val_labels"])
    assert val_out.dtype == np.float64 and val_out.shape == vp.shape
    # knots = scikit-learn's X_thresholds_ / y_thresholds_: exact integer pooling -> same knots, values to an ulp
    np.testing.assert_allclose(cal.calibrator.X_thresholds_, g["x_thresholds"], rtol=1e-15, atol=0)
    np.testing.assert_allclose(cal.calibrator.y_thresholds_, g["y_thresholds"], rtol=1e-13, atol=1e-16)
    # outputs: exp() differs by an ulp between CUDA and numpy; a steep segment of the fit amplifies that
    np.testing.assert_allclose(val_out[g["rows_val"]], g["val_out"], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(cal.transform(tp)[g["rows_test"]], g["test_out"], rtol=1e-8, atol=1e-12)
    # one-hot label matrix instead of class indices
    onehot = (d["val_labels"][:, None] == np.arange(vp.shape[1])[None]).astype(np.float64)
    cal2 = MultiIsotonicRegression()
    np.testing.assert_allclose(cal2.fit_transform(vp, onehot), val_out, rtol=0, atol=0)


def test_isotonic_fit_on_dense_softmax_tails(cuda_lib):
    """Exp-normalised logits: thousands of tail values closer than 1e-15 to their neighbours.  scikit-learn starts a
    new x value where x - FIRST x of the current value >= 1e-15 (anchored), not where neighbours differ by that
    much; the knots and the transform must follow it (79 knots against sklearn's 96 with the chained rule)."""
    from sklearn.isotonic import IsotonicRegression
    rng = np.random.default_rng(5)
    n, c = 4000, 100
    logits = rng.standard_normal((n, c)) * 9.0
    labels = torch.from_numpy(rng.integers(0, c, n))
    probs, onehot = native.exp_normalise_rows(torch.from_numpy(logits).cuda(), labels.cuda())
    x = probs.cpu().numpy().ravel()
    y = onehot.cpu().numpy().ravel()
    xs = np.sort(x)
    assert ((xs[1:] - xs[:-1]) < 1e-15).sum() > 10000, "the case must contain dense tails"
    iso = IsotonicRegression(out_of_bounds="clip").fit(x, y.astype(np.float64))
    kx, ky = native.isotonic_fit_binary(probs.reshape(-1), onehot.reshape(-1))
    np.testing.assert_allclose(kx.cpu().numpy(), iso.X_thresholds_, rtol=0, atol=0)
    np.testing.assert_allclose(ky.cpu().numpy(), iso.y_thresholds_, rtol=1e-13, atol=1e-16)
    t = rng.random(5000)
    got = native.isotonic_transform(kx, ky, torch.from_numpy(t).cuda()).cpu().numpy()
    np.testing.assert_allclose(got, iso.predict(t), rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("n,c,seed", [(1, 2, 0), (50, 3, 1), (400, 10, 2), (3000, 100, 3), (20000, 120, 4)])
def test_isotonic_fit_against_sklearn(cuda_lib, n, c, seed):
    """ccal_isotonic_fit_binary + ccal_isotonic_transform against scikit-learn on random problems with many exact
    duplicates (quantised x), all-equal targets and a single sample."""
    from sklearn.isotonic import IsotonicRegression
    rng = np.random.default_rng(seed)
    x = rng.random(n * c)
    if seed % 2 == 0:
        x = np.round(x * 50) / 50                                       # heavy ties in x
    y = (rng.random(n * c) < x).astype(np.uint8) if seed != 1 else np.ones(n * c, np.uint8)
    iso = IsotonicRegression(out_of_bounds="clip").fit(x, y.astype(np.float64))
    kx, ky = native.isotonic_fit_binary(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda())
    np.testing.assert_allclose(kx.cpu().numpy(), iso.X_thresholds_, rtol=0, atol=0)
    np.testing.assert_allclose(ky.cpu().numpy(), iso.y_thresholds_, rtol=1e-13, atol=1e-16)
    t = np.concatenate([rng.random(1000) * 1.4 - 0.2, x[:100]])        # includes out-of-range points (clipped)
    got = native.isotonic_transform(kx, ky, torch.from_numpy(t).cuda()).cpu().numpy()
    np.testing.assert_allclose(got, iso.predict(t), rtol=1e-12, atol=1e-15)


def test_bin_mean_shift_matches_reference_fixture(cuda_lib, golden):
    from clip_calibration_b200.trainers.calibration.multi_isotonic_regression import MultiIsotonicRegression
    from clip_calibration_b200.trainers.calibration.multi_proximity_isotonic import BinMeanShift
    g, d = golden("isotonic"), golden("density_ratio")
    vp, tp = d["val_probs"].astype(np.float64), d["test_probs"].astype(np.float64)
    for strategy in ("quantile", "uniform"):
        bms = BinMeanShift("multi_isotonic_regression", MultiIsotonicRegression, bin_strategy=strategy,
                           normalize_conf=False, proximity_bin=5)
        val_out = bms.fit_transform(vp, d["val_prox"], d["val_labels"])
        np.testing.assert_allclose(bms.bin_edges, g[f"bms_{strategy}_edges"], rtol=1e-7)     # float32 proximities
        np.testing.assert_allclose(val_out[g["rows_val"]], g[f"bms_{strategy}_val_out"], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(bms.transform(tp, d["test_prox"])[g["rows_test"]], g[f"bms_{strategy}_test_out"],
                                   rtol=1e-8, atol=1e-12)
    with pytest.raises(NotImplementedError):
        BinMeanShift("multi_isotonic_regression", MultiIsotonicRegression, bin_strategy="kmeans")


def test_vl_calibration_bin_based_multi_isotonic(cuda_lib, golden):
    """VLCalibration(base_calibration_mode='bin_based', base_bin_calibrator_name='multi_isotonic_regression') with and
    without proximity (reference vl_calibrator.py:97-102, :133-134, :146-148)."""
    g, d = golden("isotonic"), golden("density_ratio")
    val_logits = np.log(d["val_probs"].astype(np.float64))
    test_logits = np.log(d["test_probs"].astype(np.float64))
    val_dict = {"val_logits": val_logits, "val_labels": d["val_labels"], "val_image_features": None,
                "val_text_features": None, "val_image_knn_dists": -np.log(d["val_prox"].astype(np.float64))[:, None]}
    for procal, key in ((False, "test_out"), (True, "bms_quantile_test_out")):
        cal = vl_calibrator.VLCalibration(None, base_calibration_mode="bin_based",
                                          base_bin_calibrator_name="multi_isotonic_regression", procal_flag=procal,
                                          val_dict=val_dict)
        cal.fit()
        out = cal.predict(test_logits, d["test_prox"])
        assert out.dtype == np.float64
        # probabilities re-derived through log / fp32 softmax: knots and inputs move by ~1e-7, outputs stay close
        # except where a point crosses a step of the fitted function
        diff = np.abs(out[g["rows_test"]] - g[key])
        assert np.quantile(diff, 0.99) < 1e-4 and diff.mean() < 1e-4
    # a name outside the reference's if/elif chain builds nothing there either (vl_calibrator.py:121-148)
    none = vl_calibrator.VLCalibration(None, base_calibration_mode="bin_based", base_bin_calibrator_name="platt",
                                       val_dict=val_dict)
    none.fit()
    assert none.base_calibrator is None


def test_device_sort_and_prefix_sum(cuda_lib):
    """csrc/sort_scan.cuh (the library's own radix sort / scan, which the isotonic fit runs on): stable order of
    (float64 key, uint8 payload) pairs incl. negative keys, +-0, infinities, heavy duplicates and sizes around the
    2048-key tile; int32 prefix sums around the 4096-element tile, inclusive / exclusive / in place."""
    rng = np.random.default_rng(0)
    for n in (1, 2, 31, 2047, 2048, 2049, 4097, 100_003, 1_500_000):
        kinds = [rng.random(n), rng.standard_normal(n) * 10.0 ** rng.integers(-300, 300, n),
                 rng.integers(0, 7, n).astype(np.float64) / 7.0, np.full(n, 0.25)]
        special = rng.standard_normal(n)
        special[rng.integers(0, n, max(1, n // 50))] = 0.0
        special[rng.integers(0, n, max(1, n // 50))] = -0.0
        special[rng.integers(0, n, max(1, n // 100))] = np.inf
        special[rng.integers(0, n, max(1, n // 100))] = -np.inf
        kinds.append(special)
        for keys in kinds:
            vals = rng.integers(0, 256, n).astype(np.uint8)
            ko, vo = native.sort_pairs_f64_u8(torch.from_numpy(keys).cuda(), torch.from_numpy(vals).cuda())
            # reference order: by the key's total order (-0.0 before +0.0), ties by input position
            order = np.lexsort((np.arange(n), np.signbit(keys) == 0, keys))
            assert np.array_equal(ko.cpu().numpy().view(np.uint64), keys[order].view(np.uint64))
            assert np.array_equal(vo.cpu().numpy(), vals[order])
    for n in (1, 255, 4095, 4096, 4097, 1_000_001):
        x = rng.integers(0, 3, n).astype(np.int32)
        xd = torch.from_numpy(x).cuda()
        assert np.array_equal(native.prefix_sum_i32(xd, True).cpu().numpy(), np.cumsum(x, dtype=np.int64).astype(np.int32))
        assert np.array_equal(native.prefix_sum_i32(xd, False).cpu().numpy(),
                              (np.cumsum(x, dtype=np.int64) - x).astype(np.int32))


def test_netcal_style_calibrators_match_fixture(cuda_lib, golden):
    """HistogramBinning / IsotonicRegression (netcal's one-vs-all scheme, restated: parity with netcal itself unpinned)
    plain (vl_calibrator.py:137-143) and under the reference's BinMeanShift (:125-131): the fixture ran the reference's
    unmodified BinMeanShift around the oracle's statement of the two netcal classes."""
    from clip_calibration_b200.trainers.calibration.multi_proximity_isotonic import BinMeanShift
    from clip_calibration_b200.trainers.calibration.netcal_binning import HistogramBinning, IsotonicRegression
    g, d = golden("netcal_binning"), golden("density_ratio")
    vp, tp = d["val_probs"].astype(np.float64), d["test_probs"].astype(np.float64)
    for name, cls, kw in (("histogram_binning", HistogramBinning, {"bins": 10}), ("isotonic_regression", IsotonicRegression, {})):
        plain = cls(**kw)
        assert plain.fit(vp, d["val_labels"]) is plain
        np.testing.assert_allclose(plain.transform(vp)[g["rows_val"]], g[f"{name}_val_out"], rtol=1e-12, atol=1e-15)
        test_out = plain.transform(tp)
        assert test_out.dtype == np.float64
        np.testing.assert_allclose(test_out[g["rows_test"]], g[f"{name}_test_out"], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(test_out.sum(1), 1.0, rtol=1e-12)
        # float32 probabilities are widened exactly: same result as their float64 copy; torch in -> torch out
        out32 = plain.transform(torch.from_numpy(d["test_probs"]).cuda())
        assert out32.is_cuda and np.array_equal(out32.cpu().numpy(), test_out)
        bms = BinMeanShift(name, cls, bin_strategy="quantile", normalize_conf=False, proximity_bin=5, **kw)
        val_out = bms.fit_transform(vp, d["val_prox"], d["val_labels"])
        np.testing.assert_allclose(bms.bin_edges, g[f"bms_{name}_edges"], rtol=1e-7)
        # BinMeanShift exp-normalises its input for these two methods (reference :219-220): inputs move by an ulp,
        # the bin maps / knots with them
        np.testing.assert_allclose(val_out[g["rows_val"]], g[f"bms_{name}_val_out"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(bms.transform(tp, d["test_prox"])[g["rows_test"]], g[f"bms_{name}_test_out"],
                                   rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("n,c,seed", [(4000, 37, 0), (20_000, 101, 1), (513, 3, 2)])
def test_netcal_style_calibrators_against_the_oracle(cuda_lib, n, c, seed):
    """Random softmax outputs with classes that never occur in y, confidences exactly on bin edges (0, 0.1 ... 1.0),
    empty bins, the binary (1-D) form and independent_probabilities - against oracle.cpu_oracle.Netcal*CC."""
    from clip_calibration_b200.trainers.calibration.netcal_binning import HistogramBinning, IsotonicRegression
    rng = np.random.default_rng(seed)
    logits = rng.standard_normal((n, c)) * 3.0
    y = rng.integers(0, c - 1, n)                               # the last class never occurs
    logits[np.arange(n), y] += 2.0
    p = np.exp(logits - logits.max(1, keepdims=True))
    p /= p.sum(1, keepdims=True)
    p[:: 97, 0] = np.linspace(0.0, 1.0, 11)[rng.integers(0, 11, len(p[:: 97]))]      # values on the bin edges
    q = rng.random((n // 2, c))
    q /= q.sum(1, keepdims=True)
    for mine, ref in ((HistogramBinning(bins=15), orc.NetcalHistogramBinningCC(bins=15)),
                      (IsotonicRegression(), orc.NetcalIsotonicRegressionCC()),
                      (HistogramBinning(independent_probabilities=True), orc.NetcalHistogramBinningCC(independent_probabilities=True))):
        got = mine.fit_transform(p, y)
        want = ref.fit_transform(p, y)
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-15)
        assert np.all(got[:, c - 1] == 0.0)
        np.testing.assert_allclose(mine.transform(q), ref.transform(q), rtol=1e-12, atol=1e-15)
    conf, hit = p[:, 0], (y == 0).astype(np.int64)
    for mine, ref in ((HistogramBinning(bins=10), orc.NetcalHistogramBinningCC(bins=10)),
                      (IsotonicRegression(), orc.NetcalIsotonicRegressionCC())):
        got = mine.fit_transform(conf, hit)
        assert got.shape == (n,)
        np.testing.assert_allclose(got, ref.fit_transform(conf, hit), rtol=1e-12, atol=1e-15)
    with pytest.raises(NotImplementedError):
        HistogramBinning().fit(p[:, :2], y)
    with pytest.raises(NotImplementedError):
        HistogramBinning(equal_intervals=False)


def test_vl_calibration_bin_based_netcal_names(cuda_lib, golden):
    """VLCalibration(base_calibration_mode='bin_based') with the two netcal names, with and without proximity."""
    g, d = golden("netcal_binning"), golden("density_ratio")
    val_logits = np.log(d["val_probs"].astype(np.float64))
    test_logits = np.log(d["test_probs"].astype(np.float64))
    val_dict = {"val_logits": val_logits, "val_labels": d["val_labels"], "val_image_features": None,
                "val_text_features": None, "val_image_knn_dists": -np.log(d["val_prox"].astype(np.float64))[:, None]}
    for name in ("histogram_binning", "isotonic_regression"):
        for procal, key in ((False, f"{name}_test_out"), (True, f"bms_{name}_test_out")):
            cal = vl_calibrator.VLCalibration(None, base_calibration_mode="bin_based", base_bin_calibrator_name=name,
                                              procal_flag=procal, val_dict=val_dict)
            cal.fit()
            out = cal.predict(test_logits, d["test_prox"])
            assert out.dtype == np.float64 and out.shape == test_logits.shape
            # probabilities re-derived through log / fp32 softmax move by ~1e-7: outputs stay close except where a
            # point crosses a bin edge / a step of the fitted function
            diff = np.abs(out[g["rows_test"]] - g[key])
            assert np.quantile(diff, 0.99) < 1e-3 and diff.mean() < 1e-3, (name, procal, diff.max())


def test_custom_clip_calibration_forward_confidence(cuda_lib, golden, synth_case):
    """a-3: CustomCLIPCalibration.forward keeps the reference contract (logits, image_features, text_features);
    forward_confidence is the fused route (no [batch, C] matrix) and must agree with the oracle chain at the
    learner's current scale, with and without DAC multipliers."""
    g, case = golden("sun397_l14"), synth_case("sun397_l14")
    img = torch.from_numpy(case.img[:3000]).cuda().to(torch.bfloat16)
    txt = torch.from_numpy(case.txt_tuned).cuda().to(torch.bfloat16)

    class Base(torch.nn.Module):
        dtype = torch.bfloat16

        def forward(self, image):
            return None, image, txt

    model = tempscaling.CustomCLIPCalibration(None, Base()).cuda()
    with torch.no_grad():
        model.scale_learner.logit_scale.fill_(float(np.log(80.0)))
    scale = float(model.scale_learner().float())
    logits, f_img, f_txt = model(img)
    assert logits.shape == (3000, txt.shape[0]) and f_img is img and f_txt is txt
    cc = np.asarray(g["cc_k5"], np.float32)
    for ccx in (None, cc):
        pred, conf = model.forward_confidence(img, None if ccx is None else torch.from_numpy(ccx).cuda())
        pref, cref, gap = orc.score_chain(case.img[:3000], case.txt_tuned, ccx, scale)
        ok = gap > TIE_GAP
        assert np.array_equal(pred.cpu().numpy()[ok], pref[ok])
        np.testing.assert_allclose(conf.cpu().numpy()[ok], cref[ok], rtol=1e-4)
    # the materialised logits of forward() are a bf16 matrix (the reference contract: the base model's dtype), whose
    # rounding merges near-ties - they agree with the fp32-accumulated labels on all but a few per cent of the rows
    assert (logits.float().argmax(1).cpu().numpy()[ok] == pref[ok]).mean() > 0.9


def test_two_launch_scoring_is_bit_identical(cuda_lib, golden):
    """ccal_score_pass1 + ccal_score_pass2 == ccal_score_fused bit for bit (pred, conf, bin table), and
    from_dac(overlap_fit=True) - the DAC fit on a side stream underneath pass 1 of the head chunks - gives the same
    table and per-image outputs as the plain path."""
    case = synth.make_config("sun397_l14", seed=0)
    img = torch.from_numpy(case.img).cuda().to(torch.bfloat16)
    txt = torch.from_numpy(case.txt_tuned).cuda().to(torch.bfloat16)
    labels = torch.from_numpy(case.labels).cuda()
    cc = torch.from_numpy(np.asarray(golden("sun397_l14")["cc_k5"], dtype=np.float32)).cuda()
    thr = tm.uniform_thresholds(10)
    t1, t2 = native.new_table(10), native.new_table(10)
    pred, conf, _ = native.score_fused(img, txt, cc, case.logit_scale, labels, thr, t1)
    dotmax, pred2 = native.score_pass1(img, txt)
    conf2 = native.score_pass2(img, txt, dotmax, pred2, cc, case.logit_scale, labels, thr, t2)
    assert torch.equal(pred, pred2) and torch.equal(conf, conf2) and torch.equal(t1, t2)

    host_img = img.cpu().pin_memory()
    host_lab = labels.cpu().pin_memory()
    outs = []
    for overlap in (False, True):
        scorer = pipeline.CalibratedScorer.from_dac(case.base_zs, case.txt_zs, case.base_tuned, case.txt_tuned, k=5,
                                                    logit_scale=case.logit_scale, operand_dtype=torch.bfloat16,
                                                    group=False, overlap_fit=overlap)
        assert (scorer._fit_done is not None) == overlap
        p, c = scorer.accumulate_host(host_img, host_lab, chunk_rows=1024, keep_outputs=True)
        assert scorer._fit_done is None
        outs.append((p.cpu(), c.cpu(), scorer.reduced_table()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2], outs[1][2])
    # a scorer built with overlap_fit and used through score() first simply waits for the fit
    scorer = pipeline.CalibratedScorer.from_dac(case.base_zs, case.txt_zs, case.base_tuned, case.txt_tuned, k=5,
                                                logit_scale=case.logit_scale, operand_dtype=torch.bfloat16,
                                                group=False, overlap_fit=True)
    # reading the multipliers on the host right away must wait for the side-stream fit, not race it
    cc_now = np.asarray(scorer.dac.class_confidence)
    np.testing.assert_allclose(cc_now, golden("sun397_l14")["cc_k5"], rtol=3e-6)
    p3, c3 = scorer.score(img, labels)
    assert torch.equal(p3.cpu(), outs[0][0]) and torch.equal(c3.cpu(), outs[0][1])


def test_pipelined_evaluation_loop_equals_blocking_steps(cuda_lib):
    """An evaluation loop that queues step i+1 (text uploads on the copy stream, DAC fit on the side stream, chunked
    image uploads) before reading step i's table (reduced_table_async) returns, for every step, exactly the table of
    the blocking one-step-at-a-time form - staging buffers, the copy-stream FIFO and the allocator's stream pools
    must keep the overlapped steps apart."""
    cases = [synth.make_case(f"loop{i}", 6000 + 700 * i, 600, 300, 512, 5, 0.3, seed=10 + i) for i in range(3)]
    pin = lambda a, dt=None: (torch.from_numpy(a) if dt is None else torch.from_numpy(a).to(dt)).pin_memory()
    host = [dict(bz=pin(c.base_zs, torch.bfloat16), cz=pin(c.txt_zs, torch.bfloat16), bt=pin(c.base_tuned, torch.bfloat16),
                 ct=pin(c.txt_tuned, torch.bfloat16), img=pin(c.img, torch.bfloat16), lab=pin(c.labels)) for c in cases]

    def queue(h, overlap=True):
        scorer = pipeline.CalibratedScorer.from_dac(h["bz"], h["cz"], h["bt"], h["ct"], k=5, logit_scale=100.0,
                                                    operand_dtype=torch.bfloat16, group=False, overlap_fit=overlap)
        scorer.accumulate_host(h["img"], h["lab"], chunk_rows=1024)
        return scorer.reduced_table_async()

    blocking = [queue(h, overlap=False).result() for h in host]
    for _ in range(2):                                   # twice: the second round reuses every staging buffer
        order = [0, 1, 2, 1, 0, 2]
        pend = [queue(host[i]) for i in order]           # all six steps queued before any result is read
        for i, p in zip(order, pend):
            assert np.array_equal(p.result(), blocking[i])
            assert p.ready() and p.result() is p.result()
    assert tm.total_count(blocking[0]) == 6000


def test_guess_pipeline_predicate_matches_the_library(cuda_lib, monkeypatch):
    """native.guess_pipeline_applies (used by CalibratedScorer.score to keep large shards on the one-call path) must
    agree with what ccal_score_fused actually does, as seen through the pipeline's own row counter."""
    for n, c, d, forced in [(3000, 2048, 512, None), (3000, 2048, 512, "1"), (3000, 300, 512, "1"), (3000, 2048, 192, "1"),
                            (40_000, 30_000, 128, None), (40_000, 20_000, 128, None), (40_000, 30_000, 128, "0")]:
        if forced is None:
            monkeypatch.delenv("CCAL_SCORE_FP8", raising=False)
        else:
            monkeypatch.setenv("CCAL_SCORE_FP8", forced)
        img, txt, labels, cc = _device_case(n, c, d, 0.3, torch.bfloat16, seed=n + c + d)
        native.score_guess_stats(reset=True)
        native.score_fused(img, txt, cc, 100.0)
        rows, _ = native.score_guess_stats(reset=True)
        assert (rows == n) == native.guess_pipeline_applies(n, c, d, torch.bfloat16), (n, c, d, forced, rows)
        assert rows in (0, n)
    monkeypatch.delenv("CCAL_SCORE_FP8", raising=False)
    assert not native.guess_pipeline_applies(10 ** 6, 49408, 512, torch.float32)


@pytest.mark.parametrize("overlap", [False, True])
def test_from_dac_resolves_the_operand_dtype_from_the_callers_features(cuda_lib, overlap):
    """from_dac without operand_dtype: 16-bit text features are scored as they are (the root widens its device copy to
    fp32 for the DAC fit - inferring the operand dtype from THAT copy would give fp32 on the root and 16 bits on the
    ranks that only allocate the broadcast buffer), fp32 features take the split-precision mode; same labels either way."""
    case = synth.make_case("dtype", 3000, 640, 320, 512, 5, 0.3, seed=3, rounding=synth.round_to_fp16)
    img16 = torch.from_numpy(case.img).cuda().to(torch.float16)
    labels = torch.from_numpy(case.labels).cuda()
    preds = {}
    for name, cast in (("fp16", lambda a: torch.from_numpy(a).to(torch.float16)), ("fp32", torch.from_numpy)):
        scorer = pipeline.CalibratedScorer.from_dac(cast(case.base_zs), cast(case.txt_zs), cast(case.base_tuned),
                                                    cast(case.txt_tuned), k=5, logit_scale=100.0, group=False,
                                                    overlap_fit=overlap)
        want = torch.float16 if name == "fp16" else torch.float32
        assert scorer.operand_dtype == want and scorer.txt.dtype == want
        p, _ = scorer.score(img16 if name == "fp16" else img16.float(), labels)
        preds[name] = p.cpu()
    # the features ARE fp16 values, so both modes see the same operands; labels may differ only on near-ties
    assert (preds["fp16"] == preds["fp32"]).float().mean() > 0.999
