"""Feature hand-off formats between the reference's base-class run and its new-class run
(SURVEY.md 8f-2), so the CUDA path plugs into existing runs:

  ./temp/base_features/<dataset>/<trainer>/shots<k>/<backbone>/base/seed<s>/base_features.pt
        torch.save dict {val_logits, val_image_features, val_text_features, val_labels,
        val_image_knn_dists}        (reference trainers/classification/base_learner.py:180-184, :230-239)
  ./temp/knndist/<dataset>/<trainer>/shots<k>/<backbone>/<split>/seed<s>/nn<K>/knndist.npy
        [N_test, K] float32 distances (reference base_learner.py:123-134)

and a device-resident replacement for the evaluator's per-batch `.tolist()` accumulation
(reference evaluators/vl_evaluator.py:47-51): features stay on the GPU until they are scored.
"""
from __future__ import annotations

import os
import os.path as osp
import numpy as np
import torch

BASE_FEATURE_KEYS = ("val_logits", "val_image_features", "val_text_features", "val_labels", "val_image_knn_dists")


def base_features_path(dataset: str, trainer: str, shots: int, backbone: str, seed: int, root: str = "./temp") -> str:
    return osp.join(root, "base_features", dataset, trainer, "shots" + str(shots), backbone, "base",
                    "seed" + str(seed), "base_features.pt")


def knndist_path(dataset: str, trainer: str, shots: int, backbone: str, split: str, seed: int, k: int,
                 root: str = "./temp") -> str:
    return osp.join(root, "knndist", dataset, trainer, "shots" + str(shots), backbone, split, "seed" + str(seed),
                    "nn" + str(k), "knndist.npy")


def save_base_features(path: str, val_logits, val_image_features, val_text_features, val_labels,
                       val_image_knn_dists) -> None:
    os.makedirs(osp.dirname(path), exist_ok=True)
    to_np = lambda x: x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
    torch.save({"val_logits": to_np(val_logits), "val_image_features": to_np(val_image_features),
                "val_text_features": to_np(val_text_features), "val_labels": to_np(val_labels),
                "val_image_knn_dists": to_np(val_image_knn_dists)}, path)


def load_base_features(path: str) -> dict:
    d = torch.load(path, map_location="cpu", weights_only=False)
    missing = [k for k in BASE_FEATURE_KEYS if k not in d]
    if missing:
        raise KeyError(f"{path}: missing keys {missing}")
    return d


def load_or_compute_knn_dists(path: str, val_image_features, test_image_features, k: int) -> np.ndarray:
    """The caching rule of base_learner.py:127-134 with the CUDA kNN behind it."""
    if osp.exists(path):
        return np.load(path)
    from .trainers.calibration.proximity import get_knn_dists
    dists = get_knn_dists(val_image_features, test_image_features, k)
    os.makedirs(osp.dirname(path), exist_ok=True)
    np.save(path, dists)
    return dists


def proximity_from_knn(knn_dists) -> np.ndarray:
    """exp(-mean distance to the K nearest validation images) (base_learner.py:136-137)."""
    return np.exp(-np.mean(np.asarray(knn_dists), axis=1))


class FeatureAccumulator:
    """Collects per-batch image features and labels WITHOUT leaving the device (the reference
    converts every logit to a Python float).  `tensors()` returns the concatenated shard."""

    def __init__(self, operand_dtype=torch.bfloat16):
        self.operand_dtype = operand_dtype
        self._img, self._lab = [], []

    def process(self, image_features: torch.Tensor, labels: torch.Tensor) -> None:
        self._img.append(image_features.detach().to(self.operand_dtype))
        self._lab.append(labels.detach().to(torch.int64))

    def __len__(self) -> int:
        return sum(int(x.shape[0]) for x in self._lab)

    def tensors(self):
        return torch.cat(self._img), torch.cat(self._lab)

    def reset(self) -> None:
        self._img, self._lab = [], []
