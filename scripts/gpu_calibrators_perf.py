"""Development aid (GPU box): timings of the f-4 calibrator kernels (K6 density ratio, K7 isotonic, class counts)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_calibration_b200 import native

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
    return min(ts)


d64 = lambda *s: torch.rand(*s, device="cuda", dtype=torch.float64)
for m, n in ((4000, 50_000), (20_000, 1_000_000)):
    dx, dy, qx, qy = d64(m), 0.4 + 0.01 * d64(m), d64(n), 0.4 + 0.01 * d64(n)
    ms = timeit(lambda: native.kde2_pdf(dx, dy, qx, qy, 0.03, 0.001), reps=3, warm=1)
    print("kde2_pdf", json.dumps({"m": m, "n": n, "ms": ms, "Gpairs_s": m * n / ms / 1e6}), flush=True)

for n, c in ((50_000, 1000), (2_000_000, 100), (20_000, 49408)):
    probs = torch.softmax(torch.randn(n, c, device="cuda") * 3, dim=1)
    t, f = d64(n), d64(n)
    ms = timeit(lambda: native.density_ratio_apply(probs, t, f, 0.3))
    print("density_ratio_apply", json.dumps({"n": n, "c": c, "ms": ms, "GBs": 12.0 * n * c / ms / 1e6}), flush=True)
    labels = torch.randint(0, c, (n,), device="cuda")
    ms = timeit(lambda: native.exp_normalise_rows(probs, labels))
    print("exp_normalise_rows", json.dumps({"n": n, "c": c, "ms": ms, "GBs": 13.0 * n * c / ms / 1e6}), flush=True)
    del probs

for n in (2_000_000, 20_000_000):
    x = d64(n)
    y = (torch.rand(n, device="cuda", dtype=torch.float64) < x).to(torch.uint8)
    l0 = native.launch_count()
    kx, ky = native.isotonic_fit_binary(x, y)
    launches = native.launch_count() - l0
    ms = timeit(lambda: native.isotonic_fit_binary(x, y), reps=3, warm=1)
    print("isotonic_fit_binary", json.dumps({"n": n, "ms": ms, "knots": int(kx.numel()), "launches": launches,
                                             "Mpoints_s": n / ms / 1e3}), flush=True)
    ms = timeit(lambda: native.isotonic_transform(kx, ky, x, 1e-9))
    print("isotonic_transform", json.dumps({"n": n, "ms": ms, "GBs": 16.0 * n / ms / 1e6}), flush=True)
    del x, y

for n in (2_000_000, 20_000_000):
    x = d64(n)
    v = torch.randint(0, 2, (n,), device="cuda", dtype=torch.uint8)
    ms = timeit(lambda: native.sort_pairs_f64_u8(x, v), reps=3, warm=1)
    print("sort_pairs_f64_u8", json.dumps({"n": n, "ms": ms, "Mkeys_s": n / ms / 1e3}), flush=True)
    f = torch.randint(0, 2, (n,), device="cuda", dtype=torch.int32)
    ms = timeit(lambda: native.prefix_sum_i32(f))
    print("prefix_sum_i32", json.dumps({"n": n, "ms": ms, "GBs": 12.0 * n / ms / 1e6}), flush=True)
    del x, v, f

from clip_calibration_b200.trainers.calibration.netcal_binning import HistogramBinning, IsotonicRegression
for n, c in ((50_000, 1000), (2_000_000, 100)):
    probs = torch.softmax(torch.randn(n, c, device="cuda") * 3, dim=1)
    labels = torch.randint(0, c, (n,), device="cuda")
    edges = torch.linspace(0, 1, 11, dtype=torch.float64, device="cuda")
    ms = timeit(lambda: native.ova_hist_fit(probs, labels, edges))
    print("ova_hist_fit", json.dumps({"n": n, "c": c, "ms": ms, "GBs": 4.0 * n * c / ms / 1e6}), flush=True)
    hb = HistogramBinning(bins=10).fit_device(probs, labels)
    ms = timeit(lambda: hb.transform_device(probs))
    print("ova_apply_hist", json.dumps({"n": n, "c": c, "ms": ms, "GBs": 12.0 * n * c / ms / 1e6}), flush=True)
    if c <= 100:
        import time
        t0 = time.perf_counter(); iso = IsotonicRegression().fit_device(probs, labels); torch.cuda.synchronize()
        print("ova_isotonic_fit", json.dumps({"n": n, "c": c, "s": time.perf_counter() - t0}), flush=True)
        ms = timeit(lambda: iso.transform_device(probs))
        print("ova_apply_isotonic", json.dumps({"n": n, "c": c, "ms": ms, "GBs": 12.0 * n * c / ms / 1e6}), flush=True)
    del probs

n = 64_000_000
for c in (10, 1000, 49408):
    pred = torch.randint(0, c, (n,), device="cuda", dtype=torch.int32); gt = torch.randint(0, c, (n,), device="cuda")
    ms = timeit(lambda: native.class_counts(pred, gt, c))
    print("class_counts", json.dumps({"n": n, "c": c, "ms": ms, "GBs": 12.0 * n / ms / 1e6}), flush=True)
