"""Development aid: launch the bin-calibrator kernels once each (for ncu captures): radix sort of 20M (f64, u8) pairs,
int32 prefix sum of 20M flags, isotonic fit of 2M points, one-vs-all histogram fit / transform at 50k x 1000 and
2M x 100, one-vs-all isotonic fit + transform at 2M x 3."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_calibration_b200 import native
from clip_calibration_b200.trainers.calibration.netcal_binning import HistogramBinning, IsotonicRegression

torch.manual_seed(0)
n = 20_000_000
x = torch.rand(n, device="cuda", dtype=torch.float64)
v = torch.randint(0, 2, (n,), device="cuda", dtype=torch.uint8)
for _ in range(2):
    native.sort_pairs_f64_u8(x, v)
    native.prefix_sum_i32(v.to(torch.int32))
y = (torch.rand(2_000_000, device="cuda", dtype=torch.float64) < x[:2_000_000]).to(torch.uint8)
native.isotonic_fit_binary(x[:2_000_000].contiguous(), y)
del x, v, y
for rows, c in ((50_000, 1000), (2_000_000, 100)):
    probs = torch.softmax(torch.randn(rows, c, device="cuda") * 3, dim=1)
    labels = torch.randint(0, c, (rows,), device="cuda")
    hb = HistogramBinning(bins=10).fit_device(probs, labels)
    hb.transform_device(probs)
    del probs
probs = torch.softmax(torch.randn(2_000_000, 3, device="cuda") * 3, dim=1)      # three one-vs-all isotonic functions
labels = torch.randint(0, 3, (2_000_000,), device="cuda")
IsotonicRegression().fit_device(probs, labels).transform_device(probs)
torch.cuda.synchronize()
print("done")
