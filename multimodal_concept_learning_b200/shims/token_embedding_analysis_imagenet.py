"""Shim for ``src/multimodal/token_embedding_analysis_imagenet.py:261-286``."""
from __future__ import annotations

from typing import Dict, List

import torch

from .. import ops
from ._common import compute_device, to_kernel_dtype


def tokens_to_csr(tokenizer, token_names: List[str]):
    """Tokenise once (the reference re-tokenises per epoch, :279): CSR (offsets, ids)."""
    offsets, ids = [0], []
    for name in token_names:
        ids.extend(tokenizer.encode(name, add_special_tokens=False))
        offsets.append(len(ids))
    return torch.tensor(offsets, dtype=torch.int64), torch.tensor(ids, dtype=torch.int64)


def _gather_mean_any_device(embedding_matrix: torch.Tensor, offsets, ids, normalize=False):
    """Tables from the reference's loader live on the CPU (one [V,D] bf16 tensor per epoch).
    Only the rows that are actually referenced are uploaded (index_select = data movement);
    the mean / normalise arithmetic runs on the GPU."""
    dev = compute_device(embedding_matrix)
    if embedding_matrix.is_cuda:
        table, ids_dev = embedding_matrix, ids
    else:
        uniq, inverse = torch.unique(ids, return_inverse=True)
        if uniq.numel() == 0:
            uniq = torch.zeros(1, dtype=torch.int64)
        table, ids_dev = embedding_matrix.detach()[uniq].to(dev), inverse
    out = ops.gather_mean(to_kernel_dtype(table), offsets, ids_dev, normalize)
    return out.to(dtype=embedding_matrix.dtype, device=embedding_matrix.device)


def average_embeddings_for_tokens(tokenizer, embeddings_by_epoch: Dict[str, torch.Tensor],
                                  token_names: List[str]) -> Dict[str, torch.Tensor]:
    """Same signature and return value as the reference: ``{epoch: [n, D] table-dtype}``,
    row i = mean of the table rows of ``tokenizer.encode(token_names[i])``; empty -> zeros."""
    averaged: Dict[str, torch.Tensor] = {}
    if not embeddings_by_epoch:
        return averaged
    embedding_dim = next(iter(embeddings_by_epoch.values())).shape[1]
    offsets, ids = tokens_to_csr(tokenizer, token_names) if token_names else (None, None)
    for epoch_name, embedding_matrix in embeddings_by_epoch.items():
        if not token_names:
            averaged[epoch_name] = torch.empty((0, embedding_dim), dtype=embedding_matrix.dtype)
            continue
        averaged[epoch_name] = _gather_mean_any_device(embedding_matrix, offsets, ids)
    return averaged
