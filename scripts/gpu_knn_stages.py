"""Development aid: where does the tensor-core DAC fit spend its time?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from clip_calibration_b200 import native
torch.manual_seed(0)
def feats(n, d): return torch.nn.functional.normalize(torch.randn(n, d, device="cuda") + 1.0, dim=-1)
for b, c, d in [(1000, 49408, 512), (10000, 21841, 768)]:
    bz, cz, bt, ct = feats(b, d), feats(c, d), feats(b, d), feats(c, d)
    for name, fn in [("dac_fit", lambda: native.dac_fit(bz, cz, bt, ct, 5)), ("knn_tc", lambda: native.knn_l2(bz, cz, 5)),
                     ("knn_exhaustive", lambda: native.knn_l2(bz, cz, 5, exhaustive=True))]:
        for _ in range(3): fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        print(b, c, d, name, "ms min %.3f avg %.3f" % (min(ts), sum(ts) / len(ts)), flush=True)
