import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from clip_calibration_b200 import native
torch.manual_seed(0)
n, c, d = 1_750_000, 21841, 768
img = torch.nn.functional.normalize(torch.randn(n, d, device="cuda"), dim=-1).to(torch.bfloat16)
txt = torch.nn.functional.normalize(torch.randn(c, d, device="cuda"), dim=-1).to(torch.bfloat16)
for _ in range(2): native.score_fused(img, txt, None, 100.0, want_pred=False, want_conf=True)
ts = []
for _ in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); native.score_fused(img, txt, None, 100.0, want_pred=False, want_conf=True); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(os.environ.get("CCAL_SCORE_RESIDENT"), "in21k shard ms:", [round(t, 2) for t in ts], "TFLOP/s", round(4.0 * n * c * d / (sum(ts[2:]) / 4) / 1e9, 1), flush=True)
