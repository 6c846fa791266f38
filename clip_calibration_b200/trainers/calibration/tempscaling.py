"""Drop-in for the reference's trainers/calibration/tempscaling.py.

Kept from the reference surface (same names, parameter name `logit_scale`, init 4.6052,
checkpoint file names):
  ScaleLearner(cfg, dtype)                 reference :31-41
  CustomCLIPCalibration(cfg, base_model)   reference :44-59  (forward -> (logits, img, txt))
  TempScaling                              reference :64-327 (a dassl trainer; needs dassl)

New (the accelerated part): the objective the trainer minimises,
    L(t) = cross_entropy(exp(t) * img @ txt.T, y),   dL/dt,
evaluated from cached L2-normalised features by the fused tcgen05 kernel (ccal_ts_loss_grad)
without materialising logits, and `fit_logit_scale`, a momentum-SGD loop over that objective.

Parity note: the reference trainer's optimiser / LR schedule come from dassl's build_optimizer /
build_lr_scheduler (not in the reference tree, unpinned master).  fit_logit_scale restates the
documented defaults (SGD, momentum 0.9, weight decay 5e-4, cosine schedule, 1 constant warm-up
epoch at 1e-5, lr 0.05, 20 epochs, batch 32); the *trajectory* is therefore "parity unpinned";
the loss and gradient at a given t are pinned against torch autograd in the tests.
"""
from __future__ import annotations

import math
import os.path as osp
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from ... import native

INIT_LOG_SCALE = 4.6052          # ln(100), reference :34 and configs/calibration/TempScaling/ep20_lr5e-2.yaml


class ScaleLearner(nn.Module):
    """One learnable scalar `logit_scale`; forward() returns exp(logit_scale)."""

    def __init__(self, cfg=None, dtype=torch.float32):
        super().__init__()
        self.logit_scale = nn.Parameter(torch.tensor(INIT_LOG_SCALE, dtype=dtype))

    def forward(self):
        return self.logit_scale.exp()


class CustomCLIPCalibration(nn.Module):
    """Frozen base model + ScaleLearner.  `forward` keeps the reference contract and therefore
    returns a materialised [batch, C] logit matrix (a plain library GEMM on a 100-image batch);
    `forward_confidence` is the fused route that never builds it."""

    def __init__(self, cfg, base_model):
        super().__init__()
        self.logits_encoder = base_model
        self.dtype = base_model.dtype
        self.scale_learner = ScaleLearner(cfg, self.dtype)

    def forward(self, image, label=None):
        _, image_features, text_features = self.logits_encoder(image)
        logit_scale = self.scale_learner()
        logits = logit_scale * image_features @ text_features.t()
        return logits, image_features, text_features

    @torch.no_grad()
    def forward_confidence(self, image, class_conf=None):
        _, image_features, text_features = self.logits_encoder(image)
        dt = native.operand_dtype_for(image_features)
        pred, conf, _ = native.score_fused(image_features.to(dt).contiguous(), text_features.to(dt).contiguous(),
                                           class_conf, float(self.scale_learner().item()))
        return pred, conf


# ----------------------------------------------------------------------------------------
# objective on cached features
# ----------------------------------------------------------------------------------------
def _operands(image_features, text_features, labels, operand_dtype):
    def dev(x):
        t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
        return t.detach().cuda()
    img, txt = dev(image_features), dev(text_features)
    img = img.to(native.operand_dtype_for(img, operand_dtype))
    return img.contiguous(), txt.to(img.dtype).contiguous(), dev(labels).to(torch.int64).contiguous()


def ts_loss_and_grad(image_features, text_features, labels, log_scale: float, operand_dtype=None):
    """(loss, dloss/dlog_scale) as Python floats; one fused two-pass kernel + a fixed-order reduce."""
    img, txt, y = _operands(image_features, text_features, labels, operand_dtype)
    out = native.ts_loss_grad(img, txt, y, float(log_scale)).cpu()
    return float(out[0]), float(out[1])


def fit_logit_scale(image_features, text_features, labels, epochs: int = 20, lr: float = 0.05, batch_size: int = 32,
                    momentum: float = 0.9, weight_decay: float = 5e-4, warmup_epochs: int = 1,
                    warmup_lr: float = 1e-5, init: float = INIT_LOG_SCALE, shuffle_seed: Optional[int] = 0,
                    operand_dtype=None, return_history: bool = False):
    """Learn the scalar log-temperature on cached validation features (reference :146-169 runs
    the full CLIP forward for every batch of every epoch to fit this one parameter).

    The whole schedule runs on the device: the parameter, its momentum buffer and the running loss live in a
    4-double CUDA tensor that the loss / gradient kernel reads and ccal_sgd_scalar_step updates, the shuffled
    feature matrix is gathered once per epoch, and nothing is read back before the last step - 2 launches per
    batch and not one host synchronisation (fp16 / bf16 operands; fp32 features keep the host-stepped loop).

    Optimiser semantics - ASSUMED, not pinned: the reference builds them with dassl's `build_optimizer` /
    `build_lr_scheduler` (tempscaling.py:109-110; Dassl.pytorch is un-vendored and unpinned), driven by
    configs/calibration/TempScaling/ep20_lr5e-2.yaml and the trainer yaml's OPTIM block: `sgd`, lr 0.05, 20 epochs,
    cosine schedule, 1 warm-up epoch at constant 1e-5, batch 32 - those come from the reference tree; momentum 0.9,
    weight decay 5e-4 (applied to the scalar itself), drop_last on the train loader and "the schedule advances once
    per epoch" (`update_lr` on the last batch, tempscaling.py:166-167) are dassl defaults recalled from its source.
    The loss and its gradient at a given t ARE pinned (torch autograd, tests); the trajectory is "parity unpinned".
    """
    img, txt, y = _operands(image_features, text_features, labels, operand_dtype)
    n = img.shape[0]
    gen = torch.Generator().manual_seed(shuffle_seed) if shuffle_seed is not None else None
    on_device = img.dtype in (torch.float16, torch.bfloat16)
    state = torch.tensor([float(init), 0.0, 0.0, 0.0], dtype=torch.float64, device=img.device)
    t, vel = float(init), 0.0
    history = []
    for epoch in range(epochs):
        if epoch < warmup_epochs:
            cur_lr = warmup_lr
        else:                                   # cosine annealing over the post-warm-up epochs
            span = max(1, epochs - warmup_epochs)
            cur_lr = 0.5 * lr * (1.0 + math.cos(math.pi * (epoch - warmup_epochs) / span))
        order = torch.randperm(n, generator=gen) if gen is not None else torch.arange(n)
        order = order.to(img.device)
        img_e, y_e = img.index_select(0, order), y.index_select(0, order)        # one gather per epoch
        stops = range(0, n - batch_size + 1 if n >= batch_size else 1, batch_size)   # drop_last like dassl's train loader
        for lo in stops:
            if on_device:
                native.ts_sgd_step(img_e[lo:lo + batch_size], txt, y_e[lo:lo + batch_size], state, cur_lr, momentum,
                                   weight_decay)
            else:
                out = native.ts_loss_grad(img_e[lo:lo + batch_size], txt, y_e[lo:lo + batch_size], t).cpu()
                g = float(out[1]) + weight_decay * t
                vel = momentum * vel + g
                t -= cur_lr * vel
        if return_history and on_device:
            history.append(state.clone())       # still no synchronisation: read after the loop
    if on_device:
        t = float(state[0].item())              # the only device -> host read of the fit
    if return_history:
        hist = [h.cpu().numpy() for h in history]
        return t, [{"t": float(h[0]), "mean_loss_so_far": float(h[2] / max(h[3], 1.0))} for h in hist]
    return t


def solve_logit_scale(image_features, text_features, labels, lo: float = 0.0, hi: float = math.log(1000.0),
                      tol: float = 1e-6, max_iter: int = 60, operand_dtype=None) -> float:
    """The minimiser of the temperature-scaling objective itself (additive; no optimiser hyper-parameters):
    L(t) = CE(exp(t) * img @ txt.T, y) is convex in s = exp(t), so dL/dt has a single sign change; it is
    bracketed on [lo, hi] and located by bisection with safeguarded secant steps, each evaluation being one
    fused pass over the WHOLE cached validation set (ccal_ts_loss_grad).  Returns log-scale t*; if the
    gradient does not change sign on the bracket the better end point is returned."""
    img, txt, y = _operands(image_features, text_features, labels, operand_dtype)

    def grad(t: float) -> float:
        return float(native.ts_loss_grad(img, txt, y, t).cpu()[1])

    g_lo, g_hi = grad(lo), grad(hi)
    if g_lo >= 0.0:
        return lo
    if g_hi <= 0.0:
        return hi
    for _ in range(max_iter):
        t = hi - g_hi * (hi - lo) / (g_hi - g_lo)                  # secant (regula falsi) proposal
        if not (lo + 0.05 * (hi - lo) < t < hi - 0.05 * (hi - lo)):
            t = 0.5 * (lo + hi)                                    # safeguard: fall back to bisection
        g = grad(t)
        if g > 0.0:
            hi, g_hi = t, g
        else:
            lo, g_lo = t, g
        if hi - lo < tol or abs(g) < 1e-9:
            break
    return 0.5 * (lo + hi)


# ----------------------------------------------------------------------------------------
# checkpoint format of the scalar (dassl save_checkpoint layout, reference :260-300, :305-327)
# ----------------------------------------------------------------------------------------
def calibrated_checkpoint_name(epoch: Optional[int] = None) -> str:
    return "model-calibrated-best.pth.tar" if epoch is None else "model-calibrated.pth.tar-" + str(epoch)


def save_logit_scale(directory: str, log_scale: float, epoch: int, dtype=torch.float32, val_result=None) -> str:
    import os
    path = osp.join(directory, "tempscaling", calibrated_checkpoint_name(epoch))
    os.makedirs(osp.dirname(path), exist_ok=True)
    torch.save({"state_dict": {"logit_scale": torch.tensor(float(log_scale), dtype=dtype)}, "epoch": int(epoch),
                "optimizer": None, "scheduler": None, "val_result": val_result}, path)
    return path


def load_logit_scale(directory: str, epoch: Optional[int] = None) -> float:
    path = osp.join(directory, "tempscaling", calibrated_checkpoint_name(epoch))
    if not osp.exists(path):
        raise FileNotFoundError('Model not found at "{}"'.format(path))
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    learner = ScaleLearner()
    learner.load_state_dict(ckpt["state_dict"], strict=True)
    return float(learner.logit_scale.detach())


# ----------------------------------------------------------------------------------------
# the dassl trainer
# ----------------------------------------------------------------------------------------
class TempScaling:
    """`TempScaling` itself is trainer plumbing over the un-vendored dassl framework ("trainer
    plumbing stays in the reference") and is NOT rebuilt here.  Inside the reference repository keep the reference's
    own class (trainers/calibration/tempscaling.py:64-327) and let it build this module's CustomCLIPCalibration -
    one import line, INTEGRATION.md section 1 - or fit the scalar from cached features with fit_logit_scale() /
    solve_logit_scale() and store it with save_logit_scale(), which writes the checkpoint layout the reference's
    `load_model` reads (:260-300)."""

    def __init__(self, *a, **k):
        raise ImportError("TempScaling is a dassl trainer and stays in the reference; use the reference's class with "
                          "this module's CustomCLIPCalibration, or fit_logit_scale() on cached features")
