"""Development aid: accuracy / speed of the split-precision (fp32 operand) mode of the fused kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from clip_calibration_b200 import native, synth
from oracle import cpu_oracle as orc
ident = lambda x: np.asarray(x, np.float32)
for (n, c, d) in [(1000, 300, 512), (3000, 1000, 768), (2048, 49408, 512), (40000, 500, 128)]:
    case = synth.make_case("fp32", n, c, max(1, c // 2), d, 5, 0.3, seed=n, rounding=ident)
    cc = (0.95 + 0.05 * np.random.default_rng(1).random(c)).astype(np.float32)
    l64 = 100.0 * case.img.astype(np.float64) @ case.txt_tuned.astype(np.float64).T
    p64 = l64.argmax(1); top2 = np.partition(l64, c - 2, axis=1)[:, -2:]; gap = top2[:, 1] - top2[:, 0]
    z = cc.astype(np.float64)[p64][:, None] * (l64 - l64.max(1, keepdims=True)); c64 = 1.0 / np.exp(z).sum(1)
    for dt in (torch.float32, torch.float16, torch.bfloat16):
        img = torch.from_numpy(case.img).cuda().to(dt); txt = torch.from_numpy(case.txt_tuned).cuda().to(dt)
        pred, conf, rm = native.score_fused(img, txt, torch.from_numpy(cc).cuda(), 100.0, want_rowmax=True)
        p, cf = pred.cpu().numpy(), conf.cpu().numpy().astype(np.float64)
        ok = gap > 1e-3
        rel = np.abs(cf - c64) / c64
        print(f"n={n} c={c} d={d} {str(dt):15s} label mismatches (gap>1e-3) {(p[ok] != p64[ok]).sum():5d}/{ok.sum()}  all {(p != p64).sum():5d}  "
              f"conf rel err max {rel[p == p64].max():.2e} median {np.median(rel):.2e}  rowmax err {np.abs(rm.cpu().numpy() - l64.max(1)).max():.2e}", flush=True)
    pr, cr, _ = orc.score_chain(case.img, case.txt_tuned, cc, 100.0)
    rel = np.abs(cr - c64) / c64
    print(f"   reference fp32 path vs fp64: label mismatches {(pr != p64).sum()}  conf rel err max {rel[pr == p64].max():.2e}", flush=True)
n, c, d = 262144, 49408, 512
img = torch.nn.functional.normalize(torch.randn(n, d, device="cuda"), dim=-1); txt = torch.nn.functional.normalize(torch.randn(c, d, device="cuda"), dim=-1)
for dt in (torch.float32, torch.bfloat16):
    a, b = img.to(dt), txt.to(dt)
    for _ in range(2): native.score_fused(a, b, None, 100.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); native.score_fused(a, b, None, 100.0); e1.record(); torch.cuda.synchronize()
    print(dt, "262144 x 49408 x 512: %.2f ms" % e0.elapsed_time(e1), flush=True)
