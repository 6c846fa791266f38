"""Development aid (GPU box): ccal_ova_hist_fit against a torch statement of the same counts (every bin of every class,
label hits included), float32 and float64 inputs, class counts that need several class tiles, and its time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_calibration_b200 import native

torch.manual_seed(0)
e = torch.linspace(0, 1, 11, dtype=torch.float64, device="cuda")
for n, c, dt in ((50_000, 1000, torch.float32), (2_000_000, 100, torch.float32), (3001, 5000, torch.float64), (7, 3, torch.float32)):
    p = torch.softmax(torch.randn(n, c, device="cuda") * 3, dim=1).to(dt)
    p[::13, 0] = e[torch.randint(0, 11, (len(p[::13]),), device="cuda")].to(dt)            # values on the edges
    l = torch.randint(0, c, (n,), device="cuda")
    cnt, hit = native.ova_hist_fit(p, l, e)
    b = torch.clamp(torch.searchsorted(e, p.to(torch.float64).reshape(-1), right=True).reshape(n, c) - 1, 0, 9)
    ref = torch.zeros((c, 10), dtype=torch.int64, device="cuda")
    cls = torch.arange(c, device="cuda").expand(n, c)
    ref.index_put_((cls.reshape(-1), b.reshape(-1)), torch.ones(n * c, dtype=torch.int64, device="cuda"), accumulate=True)
    refh = torch.zeros((c, 10), dtype=torch.int64, device="cuda")
    refh.index_put_((l, b[torch.arange(n, device="cuda"), l]), torch.ones(n, dtype=torch.int64, device="cuda"), accumulate=True)
    ok = bool((ref == cnt).all()) and bool((refh == hit).all())
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0.record()
    for _ in range(5):
        native.ova_hist_fit(p, l, e)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 5
    print(f"n={n} c={c} {dt}: counts+hits == torch: {ok}; {ms:.4f} ms, {p.element_size() * n * c / ms / 1e6:.0f} GB/s", flush=True)
    assert ok
print("ok")
