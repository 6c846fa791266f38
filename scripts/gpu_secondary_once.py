"""Development aid: launch each secondary kernel a few times at a large size (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from clip_calibration_b200 import native
from clip_calibration_b200 import table_math as tm
torch.manual_seed(0)
n = 32_000_000
conf = torch.rand(n, device="cuda"); pred = torch.randint(0, 10, (n,), device="cuda", dtype=torch.int32)
gt = torch.randint(0, 10, (n,), device="cuda")
for _ in range(2):
    native.bin_stats(conf, pred, gt, tm.uniform_thresholds(10))
    native.radix_hist(conf, 0)
del conf, pred, gt
lg = torch.randn(1_000_000, 1000, device="cuda") * 5
cc = torch.ones(1000, device="cuda")
for _ in range(2):
    native.logits_confidence(lg, cc)
    native.dac_predict_logits_(lg, cc)
del lg
lg = torch.randn(20000, 49408, device="cuda") * 5
cc = torch.ones(49408, device="cuda")
for _ in range(2):
    native.logits_confidence(lg, cc)
    native.dac_predict_logits_(lg, cc)
torch.cuda.synchronize()
del lg
def feats(n, d): return torch.nn.functional.normalize(torch.randn(n, d, device="cuda") + 1.0, dim=-1)
bz, cz, bt, ct = feats(1000, 512), feats(49408, 512), feats(1000, 512), feats(49408, 512)
for _ in range(2):
    native.dac_fit(bz, cz, bt, ct, 5)
torch.cuda.synchronize()
