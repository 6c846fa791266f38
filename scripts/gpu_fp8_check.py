"""Development aid (GPU box): the FP8-guess -> bf16-verify -> redo pipeline of ccal_score_fused against the plain
two-pass kernel on the same inputs - labels must be identical, confidences equal to a few ulp, tables equal except
for samples within an ulp of a bin edge - and their timings at the bench shapes.

    python scripts/gpu_fp8_check.py            # parity cases + timing, JSON lines to gpurun_out/fp8_check.jsonl
    python scripts/gpu_fp8_check.py --one      # two pipeline calls at the open-vocabulary shape (for an ncu launch list)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from clip_calibration_b200 import native
from clip_calibration_b200 import table_math as tm

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
os.makedirs(OUT, exist_ok=True)
LOG = open(os.path.join(OUT, "fp8_check.jsonl"), "a")


def emit(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    LOG.write(line + "\n")
    LOG.flush()


def make(n, c, d, signal, n_base, dtype=torch.bfloat16, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    unit = lambda x: x / x.norm(dim=-1, keepdim=True)
    u = unit(torch.randn(d, device="cuda", generator=g))
    txt = unit(u[None, :] + 0.6 / d ** 0.5 * torch.randn(c, d, device="cuda", generator=g))
    txt = unit(txt + 0.1 / d ** 0.5 * torch.randn(c, d, device="cuda", generator=g)).to(dtype)
    labels = torch.randint(0, c, (n,), device="cuda", generator=g)
    img = torch.empty((n, d), dtype=dtype, device="cuda")
    for lo in range(0, n, 131072):
        hi = min(n, lo + 131072)
        raw = signal * txt[labels[lo:hi]].float() + torch.randn(hi - lo, d, device="cuda", generator=g) / d ** 0.5
        img[lo:hi] = unit(raw).to(dtype)
    cc = 0.984 + 0.01 * torch.rand(c, device="cuda", generator=g)
    cc[:n_base] = 1.0
    return img, txt.contiguous(), labels, cc.float().contiguous()


def run(mode, img, txt, labels, cc, thr):
    os.environ["CCAL_SCORE_FP8"] = mode
    table = native.new_table(len(thr))
    pred, conf, rowmax = native.score_fused(img, txt, cc, 100.0, labels, thr, table, want_rowmax=True)
    torch.cuda.synchronize()
    return pred, conf, rowmax, table


def timed(mode, img, txt, labels, cc, thr, reps=5):
    os.environ["CCAL_SCORE_FP8"] = mode
    table = native.new_table(len(thr))
    for _ in range(2):
        native.score_fused(img, txt, cc, 100.0, labels, thr, table)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        native.score_fused(img, txt, cc, 100.0, labels, thr, table)
        ev[i + 1].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]


def parity(name, n, c, d, signal, n_base, dtype=torch.bfloat16):
    thr = tm.uniform_thresholds(10)
    img, txt, labels, cc = make(n, c, d, signal, n_base, dtype)
    p0, c0, r0, t0 = run("0", img, txt, labels, cc, thr)
    native.score_guess_stats(reset=True)
    p1, c1, r1, t1 = run("1", img, txt, labels, cc, thr)
    rows, redone = native.score_guess_stats(reset=True)
    rel = ((c1 - c0).abs() / c0).max().item()
    t0n, t1n = native.table_to_numpy(t0), native.table_to_numpy(t1)
    emit(case=name, n=n, c=c, d=d, dtype=str(dtype), pred_mismatch=int((p0 != p1).sum().item()), conf_rel_max=rel,
         conf_bit_equal=float((c0 == c1).float().mean().item()), rowmax_abs_max=(r1 - r0).abs().max().item(),
         count_total=[int(t0n[:, 0].sum()), int(t1n[:, 0].sum())], count_diff=int(abs(t0n[:, 0].astype("int64") - t1n[:, 0].astype("int64")).sum()),
         correct_diff=int(abs(t0n[:, 1].astype("int64") - t1n[:, 1].astype("int64")).sum()),
         ece=[float(tm.ece_from_table(t0n)), float(tm.ece_from_table(t1n))], pipeline_rows=rows, redone=redone,
         acc=float((p0.long() == labels).float().mean().item()))
    # a row's result must not depend on how the shard was cut: score the first third alone
    m = n // 3
    p2, c2, _, _ = run("1", img[:m].contiguous(), txt, labels[:m].contiguous(), cc, thr)
    emit(case=name + "/subshard", rows=m, pred_equal=bool((p2 == p1[:m]).all().item()), conf_bit_equal=bool((c2 == c1[:m]).all().item()))


def main():
    if "--one" in sys.argv:
        thr = tm.uniform_thresholds(10)
        img, txt, labels, cc = make(1_000_000, 49408, 512, 0.5, 1000)
        os.environ["CCAL_SCORE_FP8"] = "1"
        table = native.new_table(10)
        for _ in range(2):
            native.score_fused(img, txt, cc, 100.0, labels, thr, table)
        torch.cuda.synchronize()
        return
    cases = [("small-forced", 3000, 2048, 512, 0.4, 200), ("mid", 50_000, 8192, 512, 0.4, 1000),
             ("openvocab-chunk", 131_072, 49408, 512, 0.5, 1000), ("in21k-part", 60_000, 21841, 768, 0.45, 10000),
             ("ragged-d640", 33_333, 5000, 640, 0.3, 500), ("lowmargin", 40_000, 4096, 512, 0.1, 100),
             ("fp16", 30_000, 4096, 256, 0.3, 100)]
    for name, n, c, d, a, b in cases:
        try:
            parity(name, n, c, d, a, b, torch.float16 if name == "fp16" else torch.bfloat16)
        except Exception as exc:  # noqa: BLE001
            emit(case=name, error=repr(exc))
            if "CUDA" in repr(exc) or "launch" in repr(exc):
                raise
    thr = tm.uniform_thresholds(10)
    for name, n, c, d, a, b in [("openvocab", 1_000_000, 49408, 512, 0.5, 1000), ("in21k", 1_750_000, 21841, 768, 0.45, 10000)]:
        img, txt, labels, cc = make(n, c, d, a, b)
        t_old = timed("0", img, txt, labels, cc, thr)
        native.score_guess_stats(reset=True)
        native.score_trace(reset=True)
        t_new = timed("1", img, txt, labels, cc, thr)
        rows, redone = native.score_guess_stats(reset=True)
        emit(trace=name, kernels=native.score_trace(reset=True))
        flops = 2.0 * n * c * d
        emit(timing=name, two_pass_ms=t_old, guess_verify_ms=t_new, speedup=sum(t_old) / sum(t_new),
             algorithmic_tflops=[flops / (sum(t_old) / len(t_old) * 1e-3) / 1e12, flops / (sum(t_new) / len(t_new) * 1e-3) / 1e12],
             redo_rate=redone / max(rows, 1))
        del img, txt, labels, cc
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
