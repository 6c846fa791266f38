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
    assert torch.equal(pp, pred[:4096][perm]) and torch.equal(cp, conf[:4096][perm])
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
