"""Shim for ``random_experiments/multi_token_embedding/multi_token.ipynb`` cells 2-3."""
from __future__ import annotations

from typing import Dict, List

import torch

from .token_embedding_analysis_imagenet import _gather_mean_any_device, tokens_to_csr


def get_averaged_embedding(text: str, tokenizer, embedding_matrix: torch.Tensor) -> torch.Tensor:
    """Cell 2 lines 1-12: mean of the token rows of ``text`` -> [D]."""
    offsets, ids = tokens_to_csr(tokenizer, [text])
    if ids.numel() == 0:           # torch.mean of an empty gather is NaN in the notebook
        return torch.full((embedding_matrix.shape[1],), float("nan"), dtype=embedding_matrix.dtype)
    return _gather_mean_any_device(embedding_matrix, offsets, ids)[0]


def get_averaged_embeddings(texts: List[str], tokenizer, embedding_matrix: torch.Tensor,
                            normalize: bool = False) -> torch.Tensor:
    """The notebook's loop over 949 colour names as ONE launch; ``normalize=True`` also applies
    cell 3 line 16 (``x / ||x||``) in the same kernel."""
    offsets, ids = tokens_to_csr(tokenizer, texts)
    return _gather_mean_any_device(embedding_matrix, offsets, ids, normalize)


def color_embeddings(colors: Dict[str, object], tokenizer, embedding_matrix: torch.Tensor,
                     normalize: bool = True) -> Dict[str, torch.Tensor]:
    names = list(colors.keys())
    e = get_averaged_embeddings(names, tokenizer, embedding_matrix, normalize)
    return {n: e[i] for i, n in enumerate(names)}
