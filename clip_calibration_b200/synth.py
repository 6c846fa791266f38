"""Synthetic feature generator shared by the oracle, the tests and bench.py.

Recipe = SURVEY.md section 8(d): unit-norm CLIP-like text features with
text-text cosine ~0.73, "tuned" features as a small perturbation of the
zero-shot ones, image features as a noisy copy of their class's tuned text
feature. Everything is float32 and then rounded to the operand dtype (bf16 by
default) and back, so the CPU oracle and the CUDA kernels consume exactly the
same representable numbers.

Base classes are the first ceil(C/2) labels, the rule the reference's datasets
use (reference datasets/oxford_pets.py:159-169).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


def round_to_bf16(x: np.ndarray) -> np.ndarray:
    """float32 -> nearest-even bf16 -> float32, in pure numpy."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    bits = x.view(np.uint32).astype(np.uint64)
    bits = (bits + 0x7FFF + ((bits >> 16) & 1)) & 0xFFFF0000
    return bits.astype(np.uint32).view(np.float32).reshape(x.shape)


def round_to_fp16(x: np.ndarray) -> np.ndarray:
    return np.asarray(x, dtype=np.float32).astype(np.float16).astype(np.float32)


def _unit(x: np.ndarray) -> np.ndarray:
    return x / np.linalg.norm(x, axis=-1, keepdims=True)


@dataclass
class SynthCase:
    """One synthetic evaluation case (all arrays float32, operand-rounded)."""
    name: str
    img: np.ndarray          # [N, D]
    labels: np.ndarray       # [N] int64
    txt_zs: np.ndarray       # [C, D] zero-shot text features of the test vocabulary
    txt_tuned: np.ndarray    # [C, D] tuned text features of the test vocabulary
    n_base: int              # base classes = range(n_base)
    k: int
    logit_scale: float
    signal: float
    seed: int

    @property
    def base_zs(self) -> np.ndarray:
        return self.txt_zs[: self.n_base]

    @property
    def base_tuned(self) -> np.ndarray:
        return self.txt_tuned[: self.n_base]


# name -> (N, C, B, D, k, signal a); BASELINE.json configs 1-5
CONFIGS = {
    "eurosat": (8100, 10, 5, 512, 5, 0.15),
    "imagenet": (50000, 1000, 500, 512, 5, 0.25),
    "sun397_l14": (19850, 397, 199, 768, 5, 0.15),
    "openvocab": (1000000, 49408, 1000, 512, 5, 0.50),
    "in21k": (14000000, 21841, 10000, 768, 5, 0.45),
}


def make_text(C: int, D: int, seed: int, rounding=round_to_bf16):
    rng = np.random.default_rng(seed)
    u = _unit(rng.standard_normal(D).astype(np.float32))
    g = rng.standard_normal((C, D)).astype(np.float32)
    txt_zs = _unit(u[None, :] + np.float32(0.6 / math.sqrt(D)) * g).astype(np.float32)
    g2 = rng.standard_normal((C, D)).astype(np.float32)
    txt_tuned = _unit(txt_zs + np.float32(0.1 / math.sqrt(D)) * g2).astype(np.float32)
    return rounding(txt_zs), rounding(txt_tuned), rng


def make_case(name: str, N: int, C: int, B: int, D: int, k: int = 5, signal: float = 0.15,
              seed: int = 0, logit_scale: float = 100.0, rounding=round_to_bf16) -> SynthCase:
    txt_zs, txt_tuned, rng = make_text(C, D, seed, rounding)
    labels = rng.integers(0, C, size=N, dtype=np.int64)
    img = np.empty((N, D), dtype=np.float32)
    step = 65536
    for lo in range(0, N, step):
        hi = min(N, lo + step)
        g = rng.standard_normal((hi - lo, D)).astype(np.float32)
        raw = np.float32(signal) * txt_tuned[labels[lo:hi]] + g * np.float32(1.0 / math.sqrt(D))
        img[lo:hi] = _unit(raw)
    return SynthCase(name, rounding(img), labels, txt_zs, txt_tuned, B, k, float(logit_scale),
                     float(signal), seed)


def make_config(name: str, seed: int = 0, n_override: int | None = None,
                rounding=round_to_bf16) -> SynthCase:
    N, C, B, D, k, a = CONFIGS[name]
    if n_override is not None:
        N = n_override
    return make_case(name, N, C, B, D, k, a, seed, 100.0, rounding)
