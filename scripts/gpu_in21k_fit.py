"""Development aid: time the DAC fit on the bench's in21k-shaped synthetic text features and report unproven rows."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from clip_calibration_b200 import native

bench.set_workload(sys.argv[1] if len(sys.argv) > 1 else "in21k")
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
img, labels, txt_zs, txt_tuned = bench.make_device_data(seed=1000)
data = {"base_zs": txt_zs[:bench.N_BASE].contiguous(), "cur_zs": txt_zs, "base_tuned": txt_tuned[:bench.N_BASE].contiguous(), "cur_tuned": txt_tuned}
del img, labels
print({k: (tuple(v.shape), v.dtype) for k, v in data.items() if hasattr(v, "shape")})
args = [data[k] for k in ("base_zs", "cur_zs", "base_tuned", "cur_tuned")]
for it in range(3):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    l0 = native.launch_count(); e0.record()
    native.dac_fit(*args, k=bench.K_DAC)
    e1.record(); torch.cuda.synchronize()
    print("fit ms", e0.elapsed_time(e1), "launches", native.launch_count() - l0)
for name, (q, r, drop) in {"zs": (args[1], args[0], False), "tuned": (args[3], args[2], False)}.items():
    for it in range(2):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); native.knn_l2(r, q, bench.K_DAC); e1.record(); torch.cuda.synchronize()
        print(name, "knn ms", e0.elapsed_time(e1))
