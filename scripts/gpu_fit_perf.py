"""Development aid: time the DAC fit (and its two kNN problems) on the bench's synthetic text features.

    python scripts/gpu_fit_perf.py [openvocab|in21k|imagenet|sun397 ...]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from clip_calibration_b200 import native

torch.cuda.set_device(0)
for name in (sys.argv[1:] or ["openvocab", "in21k"]):
    w = bench.WORKLOADS[name]
    _, _, txt_zs, txt_tuned = bench.make_device_data(w, 1024, seed=1000)
    args = [txt_zs[:w.n_base].contiguous(), txt_zs, txt_tuned[:w.n_base].contiguous(), txt_tuned]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    times = []
    for it in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = native.launch_count(); e0.record()
        native.dac_fit(*args, k=w.k)
        e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    print(name, "fit ms", [round(t, 3) for t in times], "launches", native.launch_count() - l0, flush=True)
    for tag, (q, r) in {"zs": (args[1], args[0]), "tuned": (args[3], args[2])}.items():
        ts = []
        for it in range(4):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); native.knn_l2(r, q, w.k); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(name, tag, "knn ms", [round(t, 3) for t in ts], flush=True)
