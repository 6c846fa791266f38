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
                    operand_dtype=None) -> float:
    """Learn the scalar log-temperature on cached validation features (reference :146-169 runs
    the full CLIP forward for every batch of every epoch to fit this one parameter)."""
    img, txt, y = _operands(image_features, text_features, labels, operand_dtype)
    n = img.shape[0]
    t, vel = float(init), 0.0
    gen = torch.Generator().manual_seed(shuffle_seed) if shuffle_seed is not None else None
    for epoch in range(epochs):
        if epoch < warmup_epochs:
            cur_lr = warmup_lr
        else:                                   # cosine annealing over the post-warm-up epochs
            span = max(1, epochs - warmup_epochs)
            cur_lr = 0.5 * lr * (1.0 + math.cos(math.pi * (epoch - warmup_epochs) / span))
        order = torch.randperm(n, generator=gen) if gen is not None else torch.arange(n)
        order = order.to(img.device)
        for lo in range(0, n - batch_size + 1 if n >= batch_size else 1, batch_size):   # drop_last like dassl's train loader
            sel = order[lo:lo + batch_size]
            out = native.ts_loss_grad(img[sel].contiguous(), txt, y[sel].contiguous(), t).cpu()
            g = float(out[1]) + weight_decay * t
            vel = momentum * vel + g
            t -= cur_lr * vel
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
# the dassl trainer: only definable where dassl (un-vendored, absent from this image) exists
# ----------------------------------------------------------------------------------------
try:  # pragma: no cover - dassl is not installed in the build image
    from dassl.engine import TRAINER_REGISTRY
    from trainers.classification.base_learner import VLBaseLearner
    from trainers.calibration.basemodel_loader import get_base_model
    _HAVE_DASSL = True
except Exception:  # noqa: BLE001
    _HAVE_DASSL = False

if _HAVE_DASSL:  # pragma: no cover
    from dassl.optim import build_optimizer, build_lr_scheduler
    from dassl.data import DataManager
    from dassl.utils import load_checkpoint
    import torch.nn.functional as F

    @TRAINER_REGISTRY.register()
    class TempScaling(VLBaseLearner):
        """Same trainer plumbing as the reference; the model it builds is this module's
        CustomCLIPCalibration, so test-time scoring can use forward_confidence."""

        def check_cfg(self, cfg):
            assert cfg.TRAINER.COOP.PREC in ["fp16", "fp32", "amp"]

        def build_model(self):
            cfg = self.cfg
            base_model = get_base_model(cfg, self.dm.dataset.classnames)
            base_model = self.load_base_stat(cfg, base_model)
            self.model = CustomCLIPCalibration(cfg, base_model)
            for name, param in self.model.named_parameters():
                param.requires_grad_("scale_learner" in name)
            self.model.to(self.device)
            self.optim = build_optimizer(self.model.scale_learner, cfg.OPTIM)
            self.sched = build_lr_scheduler(self.optim, cfg.OPTIM)
            self.register_model("tempscaling", self.model.scale_learner, self.optim, self.sched)
            self.scaler = None

        def build_data_loader(self):
            dm = DataManager(self.cfg)
            self.train_loader_x = dm.val_loader          # calibration uses the validation split
            self.train_loader_u = dm.train_loader_u
            self.val_loader = dm.val_loader
            self.test_loader = dm.test_loader
            self.num_classes = dm.num_classes
            self.num_source_domains = dm.num_source_domains
            self.lab2cname = dm.lab2cname
            self.dm = dm

        def parse_batch_train(self, batch):
            return batch["img"].to(self.device), batch["label"].to(self.device)

        def forward_backward(self, batch):
            image, label = self.parse_batch_train(batch)
            logits, _, _ = self.model(image, label)
            loss = F.cross_entropy(logits, label)
            self.optim.zero_grad()
            loss.backward()
            self.optim.step()
            if (self.batch_idx + 1) == self.num_batches:
                self.update_lr()
            return {"loss": loss.item()}

        def load_base_stat(self, cfg, base_model):
            if cfg.CALIBRATION.SCALING.BASE_LEARNER == "ZeroshotCLIP":
                return base_model
            learner = cfg.CALIBRATION.SCALING.BASE_LEARNER
            sub = {"MaPLe": "MultiModalPromptLearner", "CLIP_Adapter": "adapter"}.get(learner, "prompt_learner")
            epoch = cfg.CALIBRATION.SCALING.BASE_EPOCH
            model_file = "model-best.pth.tar" if epoch is None else "model.pth.tar-" + str(epoch)
            model_path = osp.join(cfg.CALIBRATION.SCALING.BASE_DIR, sub, model_file)
            if not osp.exists(model_path):
                raise FileNotFoundError('Model not found at "{}"'.format(model_path))
            state_dict = load_checkpoint(model_path)["state_dict"]
            whole_model = learner in ("MaPLe", "PromptSRC")
            prefix = "prompt_learner." if whole_model else ""
            for fixed in ("token_prefix", "token_suffix"):
                state_dict.pop(prefix + fixed, None)
            if not whole_model:
                state_dict = {f"prompt_learner.{k}": v for k, v in state_dict.items()}
            base_model.load_state_dict(state_dict, strict=False)
            if learner == "ProDA":
                base_model.set_classifier()
            return base_model

        def load_model(self, directory, epoch=None):
            if not directory:
                print("Note that load_model() is skipped as no pretrained model is given")
                return
            for name in self.get_model_names():
                model_path = osp.join(directory, name, calibrated_checkpoint_name(epoch))
                if not osp.exists(model_path):
                    raise FileNotFoundError('Model not found at "{}"'.format(model_path))
                checkpoint = load_checkpoint(model_path)
                self._models[name].load_state_dict(checkpoint["state_dict"], strict=True)

        def after_epoch(self):
            last_epoch = (self.epoch + 1) == self.max_epoch
            freq = self.cfg.TRAIN.CHECKPOINT_FREQ
            if not self.cfg.TEST.NO_TEST and self.cfg.TEST.FINAL_MODEL == "best_val":
                curr_result = self.test(split="val")
                if curr_result > self.best_result:
                    self.best_result = curr_result
                    self.save_model(self.epoch, self.output_dir, val_result=curr_result,
                                    model_name=calibrated_checkpoint_name(None))
            if last_epoch or (freq > 0 and (self.epoch + 1) % freq == 0):
                self.save_model(self.epoch, self.output_dir, model_name=calibrated_checkpoint_name(self.epoch + 1))
else:
    class TempScaling:  # noqa: D401
        """Placeholder: the trainer class needs the un-vendored dassl framework."""

        def __init__(self, *a, **k):
            raise ImportError("TempScaling is a dassl trainer; install Dassl.pytorch and run inside the reference "
                              "repo, or use fit_logit_scale() on cached features")
