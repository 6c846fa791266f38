"""CPU: host-side logic of the product (table arithmetic, quantile edges, sharding, ABI surface).
No CUDA compute is called here."""
import ctypes
import os
import re

import numpy as np
import pytest

from clip_calibration_b200 import table_math as tm
from clip_calibration_b200 import pipeline, synth
from oracle import cpu_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_table_composition_reproduces_reference_metrics(golden):
    g = golden("metric_edge_cases")
    for name in g["names"]:
        conf, pred, gt = g[f"{name}_conf"], g[f"{name}_pred"], g[f"{name}_gt"]
        for nb in (10, 15):
            tab = orc.bin_table(conf, pred, gt, tm.uniform_thresholds(nb))
            assert abs(tm.ece_from_table(tab) - float(g[f"{name}_ece{nb}"])) < 1e-7, name
            assert abs(tm.mce_from_table(tab) - float(g[f"{name}_mce{nb}"])) < 1e-7, name
            assert tm.total_count(tab) == len(conf)
            assert abs(tm.accuracy(tab) - np.mean(pred == gt)) < 1e-15


@pytest.mark.parametrize("name", ["eurosat", "sun397_l14", "imagenet"])
def test_table_composition_on_golden_cases(name, golden, synth_case):
    g, case = golden(name), synth_case(name)
    for tag in ("dac", "nodac"):
        conf, pred = g[f"{tag}_conf"], g[f"{tag}_pred"]
        tab = orc.bin_table(conf, pred, case.labels, tm.uniform_thresholds(10))
        assert abs(tm.ece_from_table(tab) - float(g[f"{tag}_ece10"])) < 1e-7
        assert abs(tm.mce_from_table(tab) - float(g[f"{tag}_mce10"])) < 1e-7
        counts = tab[:, 0].astype(np.int64)
        folded = counts[:10].copy(); folded[9] += counts[10]
        assert np.array_equal(folded, g[f"{tag}_counts10"])
        # adaptive bins through the same table machinery
        thr = orc.quantile_edges(conf, 10)[1:-1]
        assert abs(tm.sum_of_gaps(orc.bin_table(conf, pred, case.labels, thr)) - float(g[f"{tag}_ace10"])) < 1e-7


def test_quantile_ranks_match_numpy_percentile():
    rng = np.random.default_rng(0)
    for trial in range(60):
        n = int(rng.integers(1, 4000))
        nb = int(rng.choice([2, 5, 10, 15]))
        x = rng.random(n).astype(np.float32)
        if trial % 3 == 0:
            x = (np.round(x * 10) / 10).astype(np.float32)
        xs = np.sort(x)
        for method in ("averaged_inverted_cdf", "linear", "inverted_cdf"):
            lo, hi, gamma = tm.quantile_ranks(n, nb, method)
            mine = tm.lerp_like_numpy(xs[lo], xs[hi], gamma)
            ref = np.asarray(np.percentile(x, np.linspace(0, 100, nb + 1), method=method), np.float64)
            assert np.array_equal(mine, ref), (n, nb, method)
            assert np.array_equal(tm.edges_from_order_stats(xs[lo], xs[hi], gamma),
                                  orc.quantile_edges(x, nb, method))


def test_shard_bounds_partition():
    for n in (0, 1, 7, 128, 1000003):
        for world in (1, 2, 3, 8):
            cuts = [pipeline.shard_bounds(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_bf16_rounding_matches_torch():
    import torch
    rng = np.random.default_rng(1)
    x = rng.standard_normal(100000).astype(np.float32)
    assert np.array_equal(synth.round_to_bf16(x), torch.from_numpy(x).bfloat16().float().numpy())


def test_library_exports_every_declared_symbol():
    from clip_calibration_b200 import _lib
    header = open(os.path.join(ROOT, "include", "ccal.h")).read()
    declared = set(re.findall(r"CCAL_API\s+[\w\s\*]+?\b(ccal_\w+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), "ctypes table and header disagree"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib.ccal_version.restype = ctypes.c_int
    assert lib.ccal_version() == 100


def test_no_gpu_means_loud_failure():
    import torch
    from clip_calibration_b200 import _lib, native
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.CcalError):
        native.score_fused(torch.zeros(4, 64, dtype=torch.bfloat16), torch.zeros(4, 64, dtype=torch.bfloat16))
    lib = _lib.load()
    assert lib.ccal_check_device() != 0 and _lib.last_error()
    # the calibrator wrappers have no CPU path either: every entry that would compute must raise, not fall back
    with pytest.raises(_lib.CcalError):
        native.kde2_pdf(*(torch.zeros(3, dtype=torch.float64) for _ in range(4)), 0.1, 0.1)
    with pytest.raises(_lib.CcalError):
        native.isotonic_fit_binary(torch.zeros(3, dtype=torch.float64), torch.zeros(3, dtype=torch.uint8))
    with pytest.raises(_lib.CcalError):
        native.class_counts(torch.zeros(3, dtype=torch.int32), torch.zeros(3, dtype=torch.int64), 2)
    with pytest.raises(_lib.CcalError):
        native.score_pass1(torch.zeros(4, 64, dtype=torch.bfloat16), torch.zeros(4, 64, dtype=torch.bfloat16))
    from clip_calibration_b200.trainers.calibration.multi_isotonic_regression import MultiIsotonicRegression
    from clip_calibration_b200.trainers.calibration.density_ratio_calibration import DensityRatioCalibration
    with pytest.raises((RuntimeError, AssertionError)):
        MultiIsotonicRegression().fit_transform(np.full((4, 3), 1 / 3), np.array([0, 1, 2, 0]))
    with pytest.raises((RuntimeError, AssertionError)):
        DensityRatioCalibration().fit(np.full((4, 3), 1 / 3, np.float32), np.array([0, 1, 2, 0]), np.array([0, 1, 1, 0]),
                                      np.array([0.3, 0.31, 0.32, 0.33], np.float32))
    # the new entry points of round 2: device primitives of the isotonic fit and the one-vs-all calibrators
    with pytest.raises(_lib.CcalError):
        native.sort_pairs_f64_u8(torch.zeros(3, dtype=torch.float64), torch.zeros(3, dtype=torch.uint8))
    with pytest.raises(_lib.CcalError):
        native.prefix_sum_i32(torch.zeros(3, dtype=torch.int32))
    with pytest.raises(_lib.CcalError):
        native.ova_hist_fit(torch.full((4, 3), 1 / 3), torch.zeros(4, dtype=torch.int64), torch.linspace(0, 1, 11, dtype=torch.float64))
    with pytest.raises(_lib.CcalError):
        native.ova_apply(torch.full((4, 3), 1 / 3), edges=torch.linspace(0, 1, 11, dtype=torch.float64),
                         bin_map=torch.zeros((3, 10), dtype=torch.float64))
    from clip_calibration_b200.trainers.calibration.netcal_binning import HistogramBinning, IsotonicRegression
    for cal in (HistogramBinning(bins=10), IsotonicRegression()):
        with pytest.raises((RuntimeError, AssertionError)):
            cal.fit(np.full((4, 3), 1 / 3), np.array([0, 1, 2, 0]))
    with pytest.raises(NotImplementedError):
        HistogramBinning(detection=True)
    with pytest.raises(NotImplementedError):
        HistogramBinning(equal_intervals=False)


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "clip_calibration_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "cpu_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f


def test_handoff_formats_roundtrip(tmp_path):
    import torch
    from clip_calibration_b200 import handoff
    p = handoff.base_features_path("EuroSAT", "CoOp", 16, "ViT-B/16".replace("/", "-"), 1, root=str(tmp_path))
    assert p.endswith("base_features/EuroSAT/CoOp/shots16/ViT-B-16/base/seed1/base_features.pt")
    rng = np.random.default_rng(0)
    handoff.save_base_features(p, rng.random((6, 5)), torch.rand(6, 8), rng.random((5, 8)), np.arange(6), rng.random((6, 5)))
    d = handoff.load_base_features(p)
    assert set(handoff.BASE_FEATURE_KEYS) <= set(d) and d["val_image_features"].shape == (6, 8)
    kp = handoff.knndist_path("EuroSAT", "CoOp", 16, "ViT-B-16", "new", 1, 5, root=str(tmp_path))
    os.makedirs(os.path.dirname(kp)); np.save(kp, rng.random((4, 5)).astype(np.float32))
    kd = handoff.load_or_compute_knn_dists(kp, None, None, 5)           # cached file wins, no GPU touched
    assert kd.shape == (4, 5) and handoff.proximity_from_knn(kd).shape == (4,)
    acc = handoff.FeatureAccumulator(torch.float32)
    acc.process(torch.ones(3, 4), torch.arange(3)); acc.process(torch.zeros(2, 4), torch.arange(2))
    img, lab = acc.tensors()
    assert img.shape == (5, 4) and lab.dtype == torch.int64 and len(acc) == 5


def test_reliability_bins_from_table():
    rng = np.random.default_rng(2)
    conf = np.where(rng.random(5000) < 0.1, 1.0, rng.random(5000)).astype(np.float32)
    pred, gt = rng.integers(0, 3, 5000), rng.integers(0, 3, 5000)
    rb = tm.reliability_bins(orc.bin_table(conf, pred, gt, tm.uniform_thresholds(10)))
    assert np.array_equal(rb["count"], np.histogram(conf, np.linspace(0, 1, 11))[0])
    sel = conf >= 0.9
    assert abs(rb["accuracy"][9] - np.mean(pred[sel] == gt[sel])) < 1e-12
    assert abs(rb["confidence"][9] - np.mean(conf[sel].astype(np.float64))) < 1e-9


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference runs on host cores only (the reference's own functions from oracle/_ref, else the oracle port): check the line's contract here."""
    import json, subprocess, sys
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    # "reference" when oracle/_ref holds the reference's own modules (fetched by the build step wherever the reference
    # tree exists, and shipped to the GPU box with the snapshot), "port" otherwise
    from oracle import ref_loader
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_loader.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["config"]["workload"].startswith("open-vocabulary") and d["metric"].startswith("calibrated images/sec")


def test_tempscaling_surface_without_dassl(tmp_path):
    """Reference surface that must import and work without dassl / a GPU: ScaleLearner (param name, init, exp),
    CustomCLIPCalibration.forward contract, checkpoint file names and the dassl-format scalar checkpoint."""
    import torch
    from clip_calibration_b200.trainers.calibration import tempscaling as ts
    learner = ts.ScaleLearner(None, torch.float32)
    assert list(learner.state_dict()) == ["logit_scale"] and abs(float(learner.logit_scale) - 4.6052) < 1e-6
    assert abs(float(learner().detach()) - np.exp(4.6052)) < 1e-3

    class Base(torch.nn.Module):
        dtype = torch.float32

        def forward(self, image):
            f = torch.nn.functional.normalize(image, dim=-1)
            t = torch.nn.functional.normalize(torch.eye(4, 8), dim=-1)
            return f @ t.t(), f, t

    model = ts.CustomCLIPCalibration(None, Base())
    logits, img_f, txt_f = model(torch.randn(3, 8))
    assert logits.shape == (3, 4) and torch.allclose(logits, model.scale_learner() * img_f @ txt_f.t())
    assert ts.calibrated_checkpoint_name(None) == "model-calibrated-best.pth.tar"
    assert ts.calibrated_checkpoint_name(20) == "model-calibrated.pth.tar-20"
    path = ts.save_logit_scale(str(tmp_path), 4.25, 20)
    assert path.endswith(os.path.join("tempscaling", "model-calibrated.pth.tar-20"))
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ckpt) >= {"state_dict", "epoch", "optimizer", "scheduler", "val_result"} and list(ckpt["state_dict"]) == ["logit_scale"]
    assert abs(ts.load_logit_scale(str(tmp_path), 20) - 4.25) < 1e-6
    with pytest.raises(FileNotFoundError):
        ts.load_logit_scale(str(tmp_path), 99)
    with pytest.raises(ImportError):
        ts.TempScaling()                       # the trainer itself needs dassl


def test_macro_f1_from_counts_matches_sklearn():
    import warnings
    from sklearn.metrics import f1_score
    rng = np.random.default_rng(0)
    for c, n in ((5, 100), (50, 300), (397, 2000), (3, 10)):
        gt = rng.integers(0, c, n)
        pred = np.where(rng.random(n) < 0.6, gt, rng.integers(0, c + 3, n))      # some predictions outside the label set
        counts = np.zeros((int(max(gt.max(), pred.max())) + 1, 3), np.int64)
        for p, g in zip(pred, gt):
            if p == g:
                counts[g, 0] += 1
            else:
                counts[p, 1] += 1
                counts[g, 2] += 1
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = f1_score(gt, pred, average="macro", labels=np.unique(gt))     # evaluators/vl_evaluator.py:74-79
        assert abs(tm.macro_f1_from_counts(counts) - want) < 1e-12
    assert tm.macro_f1_from_counts(np.zeros((4, 3))) == 0.0


def test_reliability_diagram_data_from_table(golden):
    """tools/plot.py:11-36 arrays from an (n+1)-bin table (here built by the oracle's bin_table)."""
    from clip_calibration_b200.tools import plot
    g = golden("metric_edge_cases")
    for name in ("saturated", "on_edges", "narrow"):
        conf, pred, gt = g[f"{name}_conf"], g[f"{name}_pred"], g[f"{name}_gt"]
        for nb in (10, 15):
            table = orc.bin_table(conf, pred, gt, tm.uniform_thresholds(nb))
            d = plot.reliability_diagram_data(None, None, None, nb, table=table)
            bins = np.linspace(0, 1, nb + 1)
            idx = np.digitize(conf, bins) - 1
            acc = np.array([np.mean(gt[idx == i] == pred[idx == i]) if np.any(idx == i) else 0 for i in range(nb)])
            cf = np.array([np.mean(conf[idx == i]) if np.any(idx == i) else 0 for i in range(nb)])
            np.testing.assert_allclose(d["bin_acc"], acc, rtol=1e-12)
            np.testing.assert_allclose(d["bin_confidences"], cf, rtol=1e-6)
            np.testing.assert_allclose(d["weights"], np.histogram(conf, bins)[0] / len(conf), rtol=1e-12)
            assert abs(d["ece"] - float(g[f"{name}_ece{nb}"])) < 1e-7


def _pava_in_rounds(ones, cnt):
    """numpy statement of what csrc/isotonic.cu does on the device: every maximal run of blocks whose means do not
    strictly increase is pooled per round (exact integer cross-multiplication), until no run is left."""
    n = len(ones)
    start = np.arange(n)
    rounds = 0
    while True:
        viol = np.zeros(len(ones), bool)
        viol[:-1] = ones[:-1] * cnt[1:] >= ones[1:] * cnt[:-1]
        if not viol.any():
            break
        head = np.ones(len(ones), bool)
        head[1:] = ~viol[:-1]
        ids = np.cumsum(head) - 1
        no, nc = np.zeros(ids[-1] + 1, np.int64), np.zeros(ids[-1] + 1, np.int64)
        np.add.at(no, ids, ones)
        np.add.at(nc, ids, cnt)
        start, ones, cnt = start[head], no, nc
        rounds += 1
    fitted = np.empty(n)
    for s, e, o, c in zip(start, np.append(start[1:], n), ones, cnt):
        fitted[s:e] = o / c
    return fitted, rounds


def _anchored_unique_flags(xs, eps=1e-15):
    """numpy statement of csrc/isotonic.cu's grouping: neighbour-rule flags, next[i] = first j > i with
    xs[j] - xs[i] >= eps, then the orbit of the flagged positions under `next` marked by pointer doubling."""
    n = len(xs)
    flag = np.ones(n, bool)
    flag[1:] = (xs[1:] - xs[:-1]) >= eps
    nxt = np.empty(n, np.int64)
    for i in range(n):                                           # the device does this by binary search
        j = i + 1
        while j < n and not (xs[j] - xs[i] >= eps):
            j += 1
        nxt[i] = j
    rounds = 0
    while True:
        src = np.nonzero(flag)[0]
        tgt = nxt[src]
        tgt = tgt[tgt < n]
        new = tgt[~flag[tgt]]
        nxt = np.where(nxt < n, np.append(nxt, n)[np.minimum(nxt, n)], n)
        rounds += 1
        if len(new) == 0:
            return flag, rounds
        flag = flag.copy()
        flag[new] = True


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_anchored_grouping_matches_sklearn_make_unique(seed):
    """Dense softmax tails: many values closer than 1e-15 to their neighbours, where scikit-learn starts a new
    value when x - FIRST x of the current value >= eps (sklearn/_isotonic.pyx::_make_unique) - chaining neighbour
    differences would merge whole tails into one value."""
    from sklearn._isotonic import _make_unique
    rng = np.random.default_rng(seed)
    x = np.concatenate([rng.random(300) * 4e-14, rng.random(200) * 3e-15, rng.random(100), np.zeros(5),
                        np.exp(-rng.random(400) * 40.0) * 1e-13])
    xs = np.sort(x)
    y = (rng.random(len(xs)) < 0.5).astype(np.float64)
    flag, rounds = _anchored_unique_flags(xs)
    ux, uy, uw = _make_unique(xs, y, np.ones_like(xs))
    assert flag.sum() == len(ux) and np.array_equal(xs[flag], ux)
    gid = np.cumsum(flag) - 1
    cnt = np.bincount(gid)
    assert np.array_equal(cnt.astype(np.float64), uw)
    neighbour = np.ones(len(xs), bool)
    neighbour[1:] = (xs[1:] - xs[:-1]) >= 1e-15
    assert neighbour.sum() < flag.sum(), "the case must exercise the difference between the two rules"
    assert rounds <= 12


@pytest.mark.parametrize("n,ties,seed", [(1, False, 0), (2, False, 1), (500, True, 2), (20000, False, 3), (200000, True, 4)])
def test_pooling_in_rounds_is_the_isotonic_fit(n, ties, seed):
    """The round-wise pooling the CUDA isotonic fit uses reaches scikit-learn's sequential PAVA solution
    (sklearn/_isotonic.pyx) - fitted values, and hence the knots, agree to the last bits - in O(log n)-ish rounds."""
    from sklearn.isotonic import IsotonicRegression
    rng = np.random.default_rng(seed)
    x = rng.random(n)
    if ties:
        x = np.round(x * 100) / 100
    y = (rng.random(n) < x).astype(np.float64)
    order = np.argsort(x, kind="stable")
    xs, ys = x[order], y[order]
    flag = np.ones(n, bool)
    flag[1:] = (xs[1:] - xs[:-1]) >= 1e-15                       # sklearn _make_unique
    gid = np.cumsum(flag) - 1
    ones, cnt = np.zeros(gid[-1] + 1, np.int64), np.zeros(gid[-1] + 1, np.int64)
    np.add.at(ones, gid, ys.astype(np.int64))
    np.add.at(cnt, gid, 1)
    fitted, rounds = _pava_in_rounds(ones, cnt)
    iso = IsotonicRegression(out_of_bounds="clip").fit(x, y)
    np.testing.assert_allclose(fitted, iso.predict(xs[flag]), rtol=1e-13, atol=1e-16)
    assert rounds <= 64


def test_bench_stall_guard_fires_only_without_beats():
    """bench.StallGuard: a timed loop that beats keeps the guard quiet; once the beats stop the callback runs with
    the place that was armed (in bench.py it prints the line from what has been measured and ends the process)."""
    import sys
    import time
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import bench
    fired = []
    guard = bench.StallGuard()
    guard.on_stall = lambda where: (fired.append(where), guard.disarm())
    guard.arm("loop A", 1.5)
    for _ in range(6):
        time.sleep(0.5)
        guard.beat()
    assert fired == []
    guard.disarm()
    time.sleep(2.5)
    assert fired == []                       # disarmed: silence is fine
    guard.arm("loop B", 1.0)
    time.sleep(3.5)
    assert fired == ["loop B"]
