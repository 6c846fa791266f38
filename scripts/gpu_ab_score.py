"""Development aid: A/B two builds of libccal on the same box - alternating timed launches of ccal_score_fused on
the bench workload.  Usage: python scripts/gpu_ab_score.py scripts/ab/libccal_prev.so clip_calibration_b200/libccal.so"""
import ctypes, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_calibration_b200 import _lib

paths = sys.argv[1:3]
libs = []
for p in paths:
    lib = ctypes.CDLL(os.path.abspath(p))
    res, args = _lib.SIGNATURES["ccal_score_fused"]
    lib.ccal_score_fused.restype, lib.ccal_score_fused.argtypes = res, args
    libs.append(lib)

torch.manual_seed(0)
n, c, d = 1_000_000, 49408, 512
unit = lambda x: x / x.norm(dim=-1, keepdim=True)
txt = unit(torch.randn(c, d, device="cuda")).to(torch.bfloat16)
img = unit(torch.randn(n, d, device="cuda")).to(torch.bfloat16)
cc = (0.97 + 0.03 * torch.rand(c, device="cuda")).float()
labels = torch.randint(0, c, (n,), device="cuda")
pred = torch.empty(n, dtype=torch.int32, device="cuda"); conf = torch.empty(n, device="cuda")
table = torch.zeros((11, 3), dtype=torch.int64, device="cuda")
thr = (ctypes.c_double * 9)(*[i / 10 for i in range(1, 10)])
stream = torch.cuda.current_stream().cuda_stream


def run(lib):
    rc = lib.ccal_score_fused(img.data_ptr(), txt.data_ptr(), cc.data_ptr(), 100.0, n, c, d, 2, pred.data_ptr(),
                              conf.data_ptr(), None, labels.data_ptr(), thr, 9, table.data_ptr(), stream)
    assert rc == 0


out = {p: [] for p in paths}
for lib in libs:
    run(lib); run(lib)
torch.cuda.synchronize()
for rnd in range(4):
    for p, lib in zip(paths, libs):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            run(lib)
        e1.record(); torch.cuda.synchronize()
        out[p].append(e0.elapsed_time(e1) / 8)
for p in paths:
    print(p, ["%.3f" % t for t in out[p]], "mean %.3f ms" % (sum(out[p]) / len(out[p])))
