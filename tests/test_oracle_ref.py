"""CPU: the reference's own modules under oracle/_ref (placed there by oracle/fetch_ref.py, which the build step runs
whenever /root/reference is present) against the oracle port and the committed goldens.  This is what makes
`bench.py --impl reference` / `cpu_baseline.kind == "reference"` the reference and not a restatement of it."""
import os

import numpy as np
import pytest

from oracle import cpu_oracle as orc
from oracle import fetch_ref, ref_loader


@pytest.fixture(scope="module")
def ref():
    if not ref_loader.available():
        fetch_ref.fetch(os.environ.get("CCAL_REFERENCE", "/root/reference"), quiet=True)
    if not ref_loader.available():
        pytest.skip("oracle/_ref absent and no reference tree to fetch it from (GPU box without a pre-fetched copy)")
    return ref_loader.load()


@pytest.mark.parametrize("name,ks", [("eurosat", (5,)), ("sun397_l14", (1, 5, 10))])
def test_reference_dac_equals_port_and_golden(ref, name, ks, golden, synth_case):
    g, case = golden(name), synth_case(name)
    for k in ks:
        dac = ref.DistanseAwareCalibration()
        dac.fit(case.base_zs, case.txt_zs, case.base_tuned, case.txt_tuned, k)
        cc_ref = np.asarray(dac.class_confidence)
        cc_port, *_ = orc.dac_fit(case.base_zs, case.txt_zs, case.base_tuned, case.txt_tuned, k)
        assert np.array_equal(cc_ref, cc_port) and np.array_equal(cc_ref, g[f"cc_k{k}"])
    logits = orc.logits_fp32(case.img[:512], case.txt_tuned, case.logit_scale)
    with ref_loader.cpu_only():
        scaled = dac.predict(logits.astype(np.float64))
    assert scaled.dtype == np.float32 and np.array_equal(scaled, orc.dac_predict(logits, cc_ref))


def test_reference_metrics_equal_port(ref, golden):
    g = golden("metric_edge_cases")
    for name in g["names"]:
        conf, pred, gt = g[f"{name}_conf"], g[f"{name}_pred"], g[f"{name}_gt"]
        for nb in (10, 15):
            assert abs(float(ref.metrics.ECE(conf, pred, gt, nb)) - orc.ece(conf, pred, gt, nb)) < 1e-12, (name, nb)
            assert abs(float(ref.metrics.MCE(conf, pred, gt, nb)) - orc.mce(conf, pred, gt, nb)) < 3e-8, (name, nb)
            assert abs(float(ref.metrics.AdaptiveECE(conf, pred, gt, nb)) - orc.adaptive_ece(conf, pred, gt, nb)) < 3e-8
            assert abs(float(ref.metrics.ECE(conf, pred, gt, nb)) - float(g[f"{name}_ece{nb}"])) < 1e-12


def test_compute_acc_bin_matches_reference(ref):
    """a-13: the (lo, hi] helper of tools/metrics.py:33-55 (unused by the reference's callers, kept for the surface)."""
    from clip_calibration_b200.tools import metrics
    rng = np.random.default_rng(3)
    conf = np.concatenate([rng.random(500), [0.0, 0.1, 0.2, 1.0]])
    pred = rng.integers(0, 5, conf.size)
    true = rng.integers(0, 5, conf.size)
    for lo, hi in [(0.0, 0.1), (0.1, 0.2), (0.35, 0.65), (0.9, 1.0), (0.999, 0.9995), (0.2, 0.2)]:
        want = ref.metrics.compute_acc_bin(lo, hi, conf, pred, true)
        got = metrics.compute_acc_bin(lo, hi, conf, pred, true)
        assert len(got) == len(want) == 3
        np.testing.assert_allclose(np.asarray(got, np.float64), np.asarray(want, np.float64), rtol=1e-13, atol=0)   # pairwise vs sequential sum


def test_product_never_imports_the_reference_copy():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "clip_calibration_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "ref_loader" not in src and "oracle._ref" not in src and "oracle/_ref" not in src, f
