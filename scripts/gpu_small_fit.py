"""Development aid: DAC-fit timings on the small BASELINE.json shapes (launch/latency bound)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_calibration_b200 import native

torch.manual_seed(0)
f = lambda n, d: torch.nn.functional.normalize(torch.randn(n, d, device="cuda") + 1.0, dim=-1)
for name, b, c, d in [("eurosat", 5, 10, 512), ("sun397", 199, 397, 768), ("imagenet", 500, 1000, 512),
                      ("openvocab", 1000, 49408, 512)]:
    bz, cz, bt, ct = f(b, d), f(c, d), f(b, d), f(c, d)
    for exhaustive in (False, True):
        ts = []
        for it in range(6):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = native.launch_count()
            e0.record()
            if exhaustive:
                native.knn_l2(bz, cz, 5, exhaustive=True); native.knn_l2(bt, ct, 5, exhaustive=True)
            else:
                native.dac_fit(bz, cz, bt, ct, 5)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(name, "exhaustive-knn-only" if exhaustive else "dac_fit", "ms min %.4f median %.4f" % (min(ts), sorted(ts)[3]),
              "launches", native.launch_count() - l0, flush=True)
