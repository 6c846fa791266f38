"""Development aid: latency of small-batch scoring with and without the column-split mode."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from clip_calibration_b200 import native
torch.manual_seed(0)
c, d = 49408, 512
txt = torch.nn.functional.normalize(torch.randn(c, d, device="cuda"), dim=-1).to(torch.bfloat16)
cc = torch.ones(c, device="cuda")
for n in (1, 100, 128, 1000, 4096, 9000, 16384):
    img = torch.nn.functional.normalize(torch.randn(n, d, device="cuda"), dim=-1).to(torch.bfloat16)
    res = {}
    for mode in ("split", "nosplit"):
        if mode == "nosplit": os.environ["CCAL_SCORE_NOSPLIT"] = "1"
        else: os.environ.pop("CCAL_SCORE_NOSPLIT", None)
        for _ in range(3): native.score_fused(img, txt, cc, 100.0)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); native.score_fused(img, txt, cc, 100.0); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        res[mode] = min(ts)
    print(f"n={n:6d} x {c} x {d}: column-split {res['split']*1e3:9.1f} us   unsplit {res['nosplit']*1e3:9.1f} us   speed-up {res['nosplit']/res['split']:.1f}x", flush=True)
