"""GPU (-m gpu): BASELINE.json's headline configuration at FULL size (1M images x 49,408-word
vocabulary, 512-d, bf16) through size-independent properties, plus an oracle spot check on a
random subsample of the same rows (the oracle cannot materialise 197 GB of logits)."""
import numpy as np
import pytest
import torch

from clip_calibration_b200 import native, pipeline
from clip_calibration_b200 import table_math as tm
from oracle import cpu_oracle as orc

pytestmark = pytest.mark.gpu


def device_case(n, c, d, signal, seed):
    """SURVEY 8(d) recipe generated on the device (bf16-rounded), for sizes numpy would take minutes on."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    unit = lambda x: x / x.norm(dim=-1, keepdim=True)
    u = unit(torch.randn(d, device="cuda", generator=g))
    txt = unit(u[None, :] + 0.6 / d ** 0.5 * torch.randn(c, d, device="cuda", generator=g))
    txt = unit(txt + 0.1 / d ** 0.5 * torch.randn(c, d, device="cuda", generator=g)).to(torch.bfloat16)
    labels = torch.randint(0, c, (n,), device="cuda", generator=g)
    img = torch.empty((n, d), dtype=torch.bfloat16, device="cuda")
    for lo in range(0, n, 131072):
        hi = min(n, lo + 131072)
        raw = signal * txt[labels[lo:hi]].float() + torch.randn(hi - lo, d, device="cuda", generator=g) / d ** 0.5
        img[lo:hi] = unit(raw).to(torch.bfloat16)
    cc = (0.98 + 0.02 * torch.rand(c, device="cuda", generator=g)).float()
    cc[:1000] = 1.0
    return img, txt, labels, cc


def test_open_vocabulary_full_size_properties(cuda_lib):
    n, c, d = 1_000_000, 49408, 512
    img, txt, labels, cc = device_case(n, c, d, 0.5, 0)
    thr = tm.uniform_thresholds(10)
    table = native.new_table(10)
    pred, conf, rowmax = native.score_fused(img, txt, cc, 100.0, labels, thr, table, want_rowmax=True)
    torch.cuda.synchronize()
    tab = native.table_to_numpy(table)
    # conservation + range
    assert tm.total_count(tab) == n
    assert int(tab[:, 1].sum()) == int((pred.long() == labels).sum())
    assert float(conf.min()) > 0.0 and float(conf.max()) <= 1.0
    assert int(pred.min()) >= 0 and int(pred.max()) < c
    # the fused histogram equals the standalone kernel on the emitted (pred, conf)
    assert np.array_equal(tab, native.table_to_numpy(native.bin_stats(conf, pred, labels, thr)))
    fx = torch.round(conf.double() * float(1 << 40)).to(torch.int64).sum()
    assert int(tab[:, 2].sum()) == int(fx)
    # determinism / idempotence: a second run gives bit-identical outputs
    table2 = native.new_table(10)
    pred2, conf2, _ = native.score_fused(img, txt, cc, 100.0, labels, thr, table2)
    assert torch.equal(pred, pred2) and torch.equal(conf, conf2) and torch.equal(table, table2)
    # shard additivity (8 image shards == what 8 GPUs would all-reduce)
    table8 = native.new_table(10)
    for r in range(8):
        lo, hi = pipeline.shard_bounds(n, r, 8)
        p8, c8, _ = native.score_fused(img[lo:hi], txt, cc, 100.0, labels[lo:hi], thr, table8)
        assert torch.equal(p8, pred[lo:hi]) and torch.equal(c8, conf[lo:hi])
    assert torch.equal(table, table8)
    # row-permutation equivariance on a slice
    perm = torch.randperm(4096, device="cuda")
    pp, cp, _ = native.score_fused(img[:4096][perm].contiguous(), txt, cc, 100.0)
    # (4096 rows run in column-split mode: same canonical summation order, so bit-identical)
    assert torch.equal(pp, pred[:4096][perm])
    assert torch.equal(cp, conf[:4096][perm])
    # the predicted class's logit is the row max: recompute s*<img, txt[pred]> in fp32 on a slice
    sl = slice(0, 65536)
    dots = (img[sl].float() * txt[pred[sl].long()].float()).sum(-1) * 100.0
    assert float((dots - rowmax[sl]).abs().max()) < 2e-3
    # oracle spot check on 1536 random rows at the full vocabulary
    rows = torch.randperm(n, device="cuda")[:1536].sort().values
    pref, cref, gap = orc.score_chain(img[rows].float().cpu().numpy(), txt.float().cpu().numpy(),
                                      cc.cpu().numpy(), 100.0)
    ok = gap > 4e-5
    assert np.array_equal(pred[rows].cpu().numpy()[ok], pref[ok])
    np.testing.assert_allclose(conf[rows].cpu().numpy()[ok], cref[ok], rtol=1e-4)


def test_in21k_shard_streaming_variant_properties(cuda_lib):
    """BASELINE.json configs[4] per-GPU shard: 1.75M images x 21,841 classes x 768-d (the D > 640 streaming
    variant of the fused kernel, CTA pairs): conservation, determinism, shard additivity, oracle spot check."""
    n, c, d = 1_750_000, 21841, 768
    img, txt, labels, cc = device_case(n, c, d, 0.45, 1)
    cc[:10000] = 1.0
    thr = tm.uniform_thresholds(15)
    table = native.new_table(15)
    pred, conf, _ = native.score_fused(img, txt, cc, 100.0, labels, thr, table)
    tab = native.table_to_numpy(table)
    assert tm.total_count(tab) == n and int(tab[:, 1].sum()) == int((pred.long() == labels).sum())
    assert np.array_equal(tab, native.table_to_numpy(native.bin_stats(conf, pred, labels, thr)))
    table4 = native.new_table(15)
    for r in range(4):
        lo, hi = pipeline.shard_bounds(n, r, 4)
        p4, c4, _ = native.score_fused(img[lo:hi], txt, cc, 100.0, labels[lo:hi], thr, table4)
        assert torch.equal(p4, pred[lo:hi]) and torch.equal(c4, conf[lo:hi])
    assert torch.equal(table, table4)
    rows = torch.randperm(n, device="cuda")[:2048].sort().values
    pref, cref, gap = orc.score_chain(img[rows].float().cpu().numpy(), txt.float().cpu().numpy(), cc.cpu().numpy(), 100.0)
    ok = gap > 4e-5
    assert np.array_equal(pred[rows].cpu().numpy()[ok], pref[ok])
    np.testing.assert_allclose(conf[rows].cpu().numpy()[ok], cref[ok], rtol=1e-4)


def test_proximity_knn_at_scale_chunked(cuda_lib):
    """f-1 at production size: 600k test images against 2,000 validation images (query chunking path of the
    tensor-core kNN, 262,144 rows per chunk) vs the exhaustive scan on a slice and the oracle on a few rows."""
    g = torch.Generator(device="cuda").manual_seed(3)
    val = torch.nn.functional.normalize(torch.randn(2000, 512, device="cuda", generator=g) + 1.5, dim=-1)
    qry = torch.nn.functional.normalize(torch.randn(600_000, 512, device="cuda", generator=g) + 1.5, dim=-1)
    d_tc, i_tc = native.knn_l2(val, qry, 5)
    for lo in (0, 262144 - 500, 524288 - 500, 600_000 - 1000):
        sl = slice(lo, lo + 1000)
        d_ex, i_ex = native.knn_l2(val, qry[sl].contiguous(), 5, exhaustive=True)
        torch.testing.assert_close(d_tc[sl], d_ex, rtol=2e-6, atol=2e-7)
        assert (i_tc[sl] != i_ex).float().mean() < 2e-3
    ref = orc.knn_dists(val.cpu().numpy(), qry[-32:].cpu().numpy(), 5)
    np.testing.assert_allclose(d_tc[-32:].cpu().numpy(), ref, rtol=3e-6, atol=3e-7)
