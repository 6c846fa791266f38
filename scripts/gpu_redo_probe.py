import os, sys, json
sys.path.insert(0, '/root/repo')
import torch
import bench
from clip_calibration_b200 import native, table_math as tm
w = bench.WORKLOADS["openvocab"]
img, labels, txt_zs, txt_tuned = bench.make_device_data(w, w.n_images, seed=1000)
txt_op = txt_tuned.to(torch.bfloat16).contiguous()
cc = native.dac_fit(txt_zs[:w.n_base].contiguous(), txt_zs, txt_tuned[:w.n_base].contiguous(), txt_tuned, 5)[0]
thr = tm.uniform_thresholds(10)
table = native.new_table(10)
def run(tag, ccx, lab=labels):
    for _ in range(2): native.score_fused(img, txt_op, ccx, 100.0, lab, thr, table)
    native.score_trace(reset=True); native.score_guess_stats(reset=True)
    for _ in range(4): native.score_fused(img, txt_op, ccx, 100.0, lab, thr, table)
    tr = native.score_trace(reset=True); st = native.score_guess_stats(reset=True)
    print(tag, {k: round(v["ms_per_launch"], 3) for k, v in tr.items()}, "redo rows/call", st[1] // 4, flush=True)
run("fitted cc", cc)
g = torch.Generator(device="cuda").manual_seed(1)
cc2 = (0.984 + 0.01 * torch.rand(w.n_classes, device="cuda", generator=g)).float(); cc2[:1000] = 1.0
run("random cc", cc2)
print("unique fitted cc values", int(torch.unique(cc).numel()), "min", float(cc.min()), "max", float(cc.max()))
# where do the redo rows sit? count mis-guess rows per 256-row tile: clustered or spread
