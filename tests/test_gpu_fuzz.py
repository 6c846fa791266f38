"""GPU (-m gpu): randomized differential test of the fused scoring kernel against the oracle
(scripts/gpu_fuzz_score.py): random shapes x operand dtypes x CTA modes x column-split on/off x DAC on/off."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [3, 11])
def test_fused_scoring_fuzz(cuda_lib, seed):
    env = {k: v for k, v in os.environ.items() if k not in ("CCAL_SCORE_CTAS", "CCAL_SCORE_NOSPLIT")}
    res = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gpu_fuzz_score.py"), str(seed), "20"],
                         capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0 and "fuzz ok" in res.stdout, res.stdout[-3000:] + res.stderr[-2000:]


def test_knn_fuzz(cuda_lib):
    """Tensor-core kNN filter (+ verification + redo) against the exhaustive scan on random problems."""
    import numpy as np
    import torch
    from clip_calibration_b200 import native
    rng = np.random.default_rng(5)
    for _ in range(14):
        nr = int(rng.choice([64, 100, 257, 1000, 4000]))
        nq = int(rng.choice([128, 500, 3000, 20000]))
        d = int(rng.choice([64, 128, 512, 768]))
        k = int(rng.choice([1, 3, 5, 10, 16]))
        drop = bool(rng.integers(0, 2)) and k < 16
        spread = float(rng.choice([0.0, 1.0, 3.0]))          # 0 = isotropic (dense neighbours), 3 = clustered
        g = torch.Generator(device="cuda").manual_seed(int(rng.integers(0, 1 << 30)))
        ref = torch.nn.functional.normalize(torch.randn(nr, d, device="cuda", generator=g) + spread, dim=-1)
        qry = torch.nn.functional.normalize(torch.randn(nq, d, device="cuda", generator=g) + spread, dim=-1)
        if drop:
            qry = ref[: min(nr, nq)].contiguous()
        d_tc, i_tc = native.knn_l2(ref, qry, k, drop)
        d_ex, i_ex = native.knn_l2(ref, qry, k, drop, exhaustive=True)
        torch.testing.assert_close(d_tc, d_ex, rtol=3e-6, atol=3e-7)
        diff = i_tc != i_ex
        assert diff.float().mean() < 2e-3, (nr, nq, d, k, drop, float(diff.float().mean()))
        if diff.any():
            assert float((d_tc[diff] - d_ex[diff]).abs().max()) <= 2e-6


def test_metrics_fuzz(cuda_lib):
    """ECE / MCE / AdaptiveECE on random confidence distributions against the oracle."""
    import numpy as np
    from clip_calibration_b200.tools import metrics
    from oracle import cpu_oracle as orc
    rng = np.random.default_rng(9)
    for _ in range(16):
        n = int(rng.choice([1, 2, 17, 1000, 50000, 190000]))
        kind = int(rng.integers(0, 4))
        if kind == 0: conf = rng.random(n)
        elif kind == 1: conf = np.where(rng.random(n) < 0.5, 1.0, rng.random(n))
        elif kind == 2: conf = np.round(rng.random(n) * 8) / 8
        else: conf = 0.5 + 0.001 * rng.random(n)
        conf = np.clip(conf, 1e-6, 1.0).astype(np.float32)
        pred, gt = rng.integers(0, 3, n), rng.integers(0, 3, n)
        nb = int(rng.choice([2, 5, 10, 15, 25]))
        assert abs(float(metrics.ECE(conf, pred, gt, nb)) - orc.ece(conf, pred, gt, nb)) < 1e-7
        assert abs(float(metrics.MCE(conf, pred, gt, nb)) - orc.mce(conf, pred, gt, nb)) < 1e-7
        assert abs(float(metrics.AdaptiveECE(conf, pred, gt, nb)) - orc.adaptive_ece(conf, pred, gt, nb)) < 1e-7, (n, kind, nb)
