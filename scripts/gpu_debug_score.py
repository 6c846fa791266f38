"""Development aid (GPU box): run the fused scoring kernel on small and large shapes, print error
statistics against the oracle and raw kernel timings.  Not part of the product or the tests."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from clip_calibration_b200 import native, synth, _lib
from clip_calibration_b200 import table_math as tm
from oracle import cpu_oracle as orc

lib = _lib.load()
print("device check:", lib.ccal_check_device(), _lib.last_error(), torch.cuda.get_device_name(0), flush=True)


def run(n, c, d, dtype=torch.bfloat16, with_cc=True, seed=0):
    rounding = synth.round_to_bf16 if dtype == torch.bfloat16 else synth.round_to_fp16
    case = synth.make_case("dbg", n, c, max(1, c // 2), d, 5, 0.3, seed=seed, rounding=rounding)
    cc = (0.95 + 0.05 * np.random.default_rng(1).random(c)).astype(np.float32) if with_cc else None
    img = torch.from_numpy(case.img).cuda().to(dtype)
    txt = torch.from_numpy(case.txt_tuned).cuda().to(dtype)
    ccd = torch.from_numpy(cc).cuda() if with_cc else None
    labels = torch.from_numpy(case.labels).cuda()
    table = native.new_table(10)
    pred, conf, rowmax = native.score_fused(img, txt, ccd, 100.0, labels, tm.uniform_thresholds(10), table, want_rowmax=True)
    torch.cuda.synchronize()
    pref, cref, gap = orc.score_chain(case.img, case.txt_tuned, cc, 100.0)
    lg = orc.logits_fp32(case.img[:64], case.txt_tuned, 100.0)
    p, cf = pred.cpu().numpy(), conf.cpu().numpy()
    ok = gap > 4e-5
    mism = int((p[ok] != pref[ok]).sum())
    rel = np.abs(cf[ok] - cref[ok]) / cref[ok]
    rm = rowmax.cpu().numpy()[:64]
    print(f"n={n} c={c} d={d} {dtype}: label mismatches {mism}/{ok.sum()}  conf rel err max {rel.max():.3e} "
          f"mean {rel.mean():.3e}  rowmax abs err {np.abs(rm - lg.max(1)).max():.3e} "
          f"ece {tm.ece_from_table(native.table_to_numpy(table)):.6f} vs {orc.ece(cref, pref, case.labels, 10):.6f}", flush=True)
    if mism:
        bad = np.where(ok & (p != pref))[0][:8]
        print("   first bad rows", bad, "got", p[bad], "want", pref[bad], "conf", cf[bad], cref[bad], flush=True)


print("CCAL_SCORE_CTAS =", os.environ.get("CCAL_SCORE_CTAS"), flush=True)
for shape in [(128, 256, 64), (128, 256, 512), (1000, 300, 512), (129, 257, 128), (300, 1000, 768), (2048, 49408, 512), (20000, 1000, 512), (20001, 397, 768)]:
    try:
        run(*shape)
    except Exception as e:  # noqa: BLE001
        print("FAILED", shape, repr(e), flush=True)
        raise
run(1000, 300, 512, torch.float16)

# timing at the headline size
n, c, d = 1_000_000, 49408, 512
g = torch.Generator(device="cuda").manual_seed(0)
txt = torch.nn.functional.normalize(torch.randn(c, d, device="cuda", generator=g), dim=-1).to(torch.bfloat16)
img = torch.nn.functional.normalize(torch.randn(n, d, device="cuda", generator=g), dim=-1).to(torch.bfloat16)
labels = torch.randint(0, c, (n,), device="cuda", generator=g)
cc = torch.ones(c, device="cuda")
table = native.new_table(10)
thr = tm.uniform_thresholds(10)
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    native.score_fused(img, txt, cc, 100.0, labels, thr, table, want_pred=True, want_conf=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"1M x 49408 x 512: {ms:.2f} ms  {n / ms * 1e3 / 1e6:.2f} M img/s  executed {4 * n * c * d / ms / 1e9:.1f} TFLOP/s", flush=True)
