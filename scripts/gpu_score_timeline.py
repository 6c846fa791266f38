"""Development aid: per-kernel CUDA-event timeline of the scoring pipeline at the bench's headline shape.

    CCAL_SCORE_TIMELINE=1 python scripts/gpu_score_timeline.py [openvocab|in21k] [rows]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from clip_calibration_b200 import native
from clip_calibration_b200 import table_math as tm

torch.cuda.set_device(0)
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "openvocab"]
n = int(sys.argv[2]) if len(sys.argv) > 2 else w.n_images
img, labels, txt_zs, txt_tuned = bench.make_device_data(w, n, seed=1000)
txt_op = txt_tuned.to(torch.bfloat16).contiguous()
cc = native.dac_fit(txt_zs[:w.n_base].contiguous(), txt_zs, txt_tuned[:w.n_base].contiguous(), txt_tuned, w.k)[0]
thr = tm.uniform_thresholds(10)
table = native.new_table(10)
for it in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    native.score_fused(img, txt_op, cc, 100.0, labels, thr, table, want_pred=True, want_conf=True)
    e1.record(); torch.cuda.synchronize()
    print("call ms", round(e0.elapsed_time(e1), 3), flush=True)
