"""Development aid: randomized differential test of the fused scoring kernel against the CPU oracle over
shapes, operand dtypes, CTA modes, DAC on/off, bin counts.  Exits non-zero on the first disagreement."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from clip_calibration_b200 import native, synth
from clip_calibration_b200 import table_math as tm
from oracle import cpu_oracle as orc

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 60
rounders = {torch.bfloat16: synth.round_to_bf16, torch.float16: synth.round_to_fp16, torch.float32: lambda x: np.asarray(x, np.float32)}
for case_i in range(n_cases):
    n = int(rng.choice([1, 7, 100, 128, 129, 300, 1000, 2500, 5000, 20000]))
    c = int(rng.choice([1, 2, 10, 255, 256, 257, 1000, 1025, 3000, 6000, 12000]))
    d = int(rng.choice([64, 128, 256, 512, 640, 768, 1024]))
    if n * c > 4e7: c = max(1, int(4e7 // n))
    dtype = [torch.bfloat16, torch.float16, torch.float32][int(rng.integers(0, 3))]
    ctas = str(rng.choice(["1", "2", ""]))
    nosplit = bool(rng.integers(0, 4) == 0)
    use_cc = bool(rng.integers(0, 2))
    nb = int(rng.choice([5, 10, 15, 20]))
    scale = float(rng.choice([1.0, 50.0, 100.0]))
    for k, v in (("CCAL_SCORE_CTAS", ctas), ("CCAL_SCORE_NOSPLIT", "1" if nosplit else "")):
        if v: os.environ[k] = v
        else: os.environ.pop(k, None)
    case = synth.make_case("fuzz", n, c, max(1, c // 2), d, 5, float(rng.choice([0.1, 0.3, 0.6])), seed=int(rng.integers(0, 1 << 30)), logit_scale=scale, rounding=rounders[dtype])
    cc = (0.9 + 0.2 * rng.random(c)).astype(np.float32) if use_cc else None
    img = torch.from_numpy(case.img).cuda().to(dtype); txt = torch.from_numpy(case.txt_tuned).cuda().to(dtype)
    thr = tm.uniform_thresholds(nb); table = native.new_table(nb)
    pred, conf, rmax = native.score_fused(img, txt, torch.from_numpy(cc).cuda() if use_cc else None, scale, torch.from_numpy(case.labels).cuda(), thr, table, want_rowmax=True)
    pred, conf = pred.cpu().numpy(), conf.cpu().numpy()
    pref, cref, gap = orc.score_chain(case.img, case.txt_tuned, cc, scale)
    ok = gap > (2e-4 if dtype == torch.float32 else 4e-5) * max(scale, 1.0) / 100.0
    tag = f"case {case_i}: n={n} c={c} d={d} {dtype} ctas={ctas or 'auto'} nosplit={nosplit} cc={use_cc} bins={nb} s={scale}"
    bad = int((pred[ok] != pref[ok]).sum())
    rel = float(np.max(np.abs(conf[ok] - cref[ok]) / cref[ok])) if ok.any() else 0.0
    tab_ok = np.array_equal(native.table_to_numpy(table), orc.bin_table(conf, pred, case.labels, thr))
    print(tag, f"-> mismatches {bad}, conf rel {rel:.2e}, table {'ok' if tab_ok else 'BAD'}", flush=True)
    if bad or rel > 1e-4 or not tab_ok:
        print("FUZZ FAILURE"); sys.exit(1)
print("fuzz ok")
