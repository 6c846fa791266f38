"""Nearest vocabulary words of learned context vectors: the search step of the
reference's interpret_prompts/interpret_prompt.py:69-72 (`torch.cdist(ctx, token_embedding)` + argsort
top-k over the 49,408-token CLIP vocabulary) on the CUDA kNN kernel (ccal_knn_l2).  Decoding the token
ids to strings stays with the reference's tokenizer."""
from __future__ import annotations

import numpy as np
import torch

from . import native


def nearest_tokens(ctx, token_embedding, topk: int = 5):
    """ctx [n_ctx, D], token_embedding [V, D] -> (token ids [n_ctx, topk] int32, distances [n_ctx, topk] float32),
    nearest first, Euclidean distance (what torch.cdist computes); topk <= 16."""
    as_numpy = not isinstance(ctx, torch.Tensor)
    dev = lambda x: (torch.from_numpy(np.ascontiguousarray(x)) if not isinstance(x, torch.Tensor) else x.detach()) \
        .to(device="cuda", dtype=torch.float32).contiguous()
    dist, idx = native.knn_l2(dev(token_embedding), dev(ctx), int(topk))
    return (idx.cpu().numpy(), dist.cpu().numpy()) if as_numpy else (idx, dist)
