"""Development aid: launch the round-2 secondary kernels once each at the bench's shapes (for ncu captures):
DAC fit at the open-vocabulary and in21k shapes with general fp32 features (hi/lo split, three MMAs) and with
bf16-valued features (single-MMA mode), then one scoring call at 262,144 x 49,408 x 512 (quantisation, persistent
guessed-class-logit kernel, redo)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from clip_calibration_b200 import native

torch.cuda.set_device(0)
for name in ("openvocab", "in21k"):
    w = bench.WORKLOADS[name]
    img, labels, txt_zs, txt_tuned = bench.make_device_data(w, 262144 if name == "openvocab" else 1024, seed=1000)
    g = torch.Generator(device="cuda").manual_seed(1)
    noise = lambda t: torch.nn.functional.normalize(t + 1e-4 * torch.randn(t.shape, device="cuda", generator=g), dim=-1)
    for tag, (zs, tu) in {"bf16-valued": (txt_zs, txt_tuned), "fp32": (noise(txt_zs), noise(txt_tuned))}.items():
        for _ in range(2):
            native.dac_fit(zs[:w.n_base].contiguous(), zs, tu[:w.n_base].contiguous(), tu, w.k)
        torch.cuda.synchronize()
    if name == "openvocab":
        cc = native.dac_fit(txt_zs[:w.n_base].contiguous(), txt_zs, txt_tuned[:w.n_base].contiguous(), txt_tuned, w.k)[0]
        for _ in range(2):
            native.score_fused(img, txt_tuned.to(torch.bfloat16).contiguous(), cc, 100.0)
        torch.cuda.synchronize()
    del img, labels, txt_zs, txt_tuned
