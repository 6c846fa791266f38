"""Development aid: how many kNN rows fail the tensor-core filter's proof (CCAL_KNN_DEBUG=1 prints the count)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["CCAL_KNN_DEBUG"] = "1"
import numpy as np, torch
from clip_calibration_b200 import native, synth
for name, C, B, D in [("recipe-openvocab", 49408, 1000, 512), ("recipe-in21k", 21841, 10000, 768), ("recipe-imagenet", 1000, 500, 512)]:
    zs, tuned, _ = synth.make_text(C, D, 0)
    zs, tuned = torch.from_numpy(zs).cuda(), torch.from_numpy(tuned).cuda()
    print(name, flush=True)
    native.dac_fit(zs[:B].contiguous(), zs, tuned[:B].contiguous(), tuned, 5)
    torch.cuda.synchronize()
g = torch.Generator(device="cuda").manual_seed(0)
val = torch.nn.functional.normalize(torch.randn(2000, 512, device="cuda", generator=g) + 1.5, dim=-1)
qry = torch.nn.functional.normalize(torch.randn(100000, 512, device="cuda", generator=g) + 1.5, dim=-1)
print("random prox", flush=True); native.knn_l2(val, qry, 5); torch.cuda.synchronize()
