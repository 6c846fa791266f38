"""Development aid: one launch of each fused-kernel variant at a representative size (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from clip_calibration_b200 import native
torch.manual_seed(0)
def feats(n, d, dt): return torch.nn.functional.normalize(torch.randn(n, d, device="cuda"), dim=-1).to(dt)
for (n, c, d, dt) in [(262144, 21841, 768, torch.bfloat16), (131072, 49408, 1024, torch.bfloat16), (131072, 49408, 512, torch.float32),
                      (50000, 1000, 512, torch.bfloat16), (262144, 49408, 512, torch.float16)]:
    img, txt = feats(n, d, dt), feats(c, d, dt)
    for _ in range(2):
        native.score_fused(img, txt, None, 100.0)
    torch.cuda.synchronize()
