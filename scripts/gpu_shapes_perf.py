"""Development aid (GPU box): kernel timings over the BASELINE.json config shapes and the HBM-bound
secondary kernels, with achieved TFLOP/s / GB/s.  Not part of the product or the tests."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from clip_calibration_b200 import native
from clip_calibration_b200 import table_math as tm

torch.manual_seed(0)


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)


def feats(n, d, dtype=torch.bfloat16):
    return torch.nn.functional.normalize(torch.randn(n, d, device="cuda"), dim=-1).to(dtype)


out = {}
thr = tm.uniform_thresholds(10)
for name, n, c, d in [("eurosat", 8100, 10, 512), ("imagenet", 50000, 1000, 512), ("sun397_l14", 19850, 397, 768),
                      ("openvocab", 1_000_000, 49408, 512), ("in21k_1gpu_shard", 1_750_000, 21841, 768),
                      ("d640", 500_000, 49408, 640), ("d1024", 250_000, 49408, 1024), ("d256", 1_000_000, 49408, 256)]:
    img, txt = feats(n, d), feats(c, d)
    labels = torch.randint(0, c, (n,), device="cuda")
    cc = torch.ones(c, device="cuda")
    table = native.new_table(10)
    mn, av = timeit(lambda: native.score_fused(img, txt, cc, 100.0, labels, thr, table, want_pred=False, want_conf=False))
    out[name] = {"n": n, "c": c, "d": d, "ms_min": mn, "ms_avg": av, "Mimg_s": n / mn / 1e3,
                 "executed_TFLOPs": 4.0 * n * c * d / mn / 1e9}
    print(name, json.dumps(out[name]), flush=True)
    del img, txt

# K1: DAC fit shapes
for name, b, c, d in [("dacfit_openvocab", 1000, 49408, 512), ("dacfit_in21k", 10000, 21841, 768), ("dacfit_imagenet", 500, 1000, 512)]:
    bz, cz, bt, ct = (feats(b, d, torch.float32), feats(c, d, torch.float32), feats(b, d, torch.float32), feats(c, d, torch.float32))
    mn, av = timeit(lambda: native.dac_fit(bz, cz, bt, ct, 5), reps=3, warm=1)
    out[name] = {"b": b, "c": c, "d": d, "ms_min": mn, "fp32_TFLOPs": 2 * 3.0 * b * c * d / mn / 1e9}
    print(name, json.dumps(out[name]), flush=True)

# proximity kNN: 100k test images vs 2000 val images
ref, q = feats(2000, 512, torch.float32), feats(100000, 512, torch.float32)
mn, av = timeit(lambda: native.knn_l2(ref, q, 5), reps=3, warm=1)
print("knn_prox", json.dumps({"ms_min": mn, "fp32_TFLOPs": 3.0 * 2000 * 100000 * 512 / mn / 1e9}), flush=True)

# K3 / K4: HBM-bound
n = 64_000_000
conf = torch.rand(n, device="cuda"); pred = torch.randint(0, 10, (n,), device="cuda", dtype=torch.int32)
gt = torch.randint(0, 10, (n,), device="cuda")
mn, av = timeit(lambda: native.bin_stats(conf, pred, gt, thr))
print("bin_stats", json.dumps({"n": n, "ms_min": mn, "GBs": 16.0 * n / mn / 1e6}), flush=True)
mn, av = timeit(lambda: native.radix_hist(conf, 0))
print("radix_hist0", json.dumps({"n": n, "ms_min": mn, "GBs": 4.0 * n / mn / 1e6}), flush=True)
del conf, pred, gt
for n, c in [(50000, 1000), (2_000_000, 1000), (40000, 49408), (20_000_000, 10)]:
    lg = torch.randn(n, c, device="cuda") * 5
    cc = torch.ones(c, device="cuda")
    mn, av = timeit(lambda: native.logits_confidence(lg, cc))
    print("logits_confidence", json.dumps({"n": n, "c": c, "ms_min": mn, "GBs": 4.0 * n * c / mn / 1e6}), flush=True)
    mn, av = timeit(lambda: native.dac_predict_logits_(lg, cc))
    print("dac_predict_logits", json.dumps({"n": n, "c": c, "ms_min": mn, "GBs": 8.0 * n * c / mn / 1e6}), flush=True)
    del lg
