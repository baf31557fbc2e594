"""Shim for ``src/multimodal/token_embedding_analysis.py`` (reference lines cited inline)."""
from __future__ import annotations

import re

import numpy as np
import torch

from .. import ops
from ._common import compute_device, to_kernel_dtype


def extract_rgb_from_mapping(labels_mapping, token):
    """Reference :173-181 (host string work, unchanged semantics)."""
    for rgb_key, token_value in labels_mapping.items():
        if token_value == token:
            match = re.match(r"r(\d+)g(\d+)b(\d+)", rgb_key)
            if match:
                r, g, b = map(int, match.groups())
                return (r / 255.0, g / 255.0, b / 255.0)
    return (0.5, 0.5, 0.5)


def pairwise_cosine_similarity(token_embeddings: torch.Tensor) -> torch.Tensor:
    """All-pairs cosine similarity [n,n] (fp32, on the GPU) -- ONE fused launch in place of the
    n(n-1)/2 ``cosine_similarity([a],[b])`` calls of reference :237-246."""
    dev = compute_device(token_embeddings)
    e = to_kernel_dtype(token_embeddings.detach()).to(dev)
    return ops.similarity_matrix(e, e, normalize=True)


def calculate_color_embedding_correlation(embeddings_by_epoch, ood_tokens, regular_tokens,
                                          ood_token_ids, regular_token_ids, labels_mapping):
    """Pearson r between RGB L1 distance and embedding cosine distance over all token pairs.
    Same signature, prints and return value as reference :183-260."""
    epoch_names = [name for name in embeddings_by_epoch.keys() if name.startswith("epoch_")]
    if not epoch_names:
        print("No epoch data found, using initial embeddings")
        last_epoch_name = "initial"
    else:
        last_epoch_name = f"epoch_{max(int(name.split('_')[1]) for name in epoch_names)}"
    print(f"\n=== Color-Embedding Distance Correlation Analysis ({last_epoch_name}) ===")

    embedding_matrix = embeddings_by_epoch[last_epoch_name]
    all_token_ids = list(ood_token_ids) + list(regular_token_ids)
    all_token_names = list(ood_tokens) + list(regular_tokens)
    n_tokens = len(all_token_names)

    # :220 gather the concept rows (data movement), then one GPU launch for every pair
    token_embeddings = embedding_matrix[all_token_ids].detach()
    print(f"Calculating pairwise distances for {n_tokens} tokens (regular + OOD)...")
    cos = pairwise_cosine_similarity(token_embeddings).cpu().numpy()

    rgb_colors = np.array([extract_rgb_from_mapping(labels_mapping, t) for t in all_token_names])
    iu = np.triu_indices(n_tokens, k=1)                      # i < j in the loop order of :237-238
    color_distances = np.abs(rgb_colors[iu[0]] - rgb_colors[iu[1]]).sum(axis=1)     # :240
    embedding_distances = 1 - cos[iu]                                                  # :245
    correlation = np.corrcoef(color_distances, embedding_distances)[0, 1]              # :253

    print(f"Number of token pairs: {len(color_distances)}")
    print(f"Color distance range: [{color_distances.min():.4f}, {color_distances.max():.4f}]")
    print(f"Embedding distance range: [{embedding_distances.min():.4f}, {embedding_distances.max():.4f}]")
    print(f"Pearson correlation coefficient: {correlation:.4f}")
    return correlation


def nearest_tokens(embedding_matrix: torch.Tensor, query_token_ids, k: int = 50,
                   inv_norm_table: torch.Tensor | None = None):
    """Cosine top-k neighbours of concept tokens in the WHOLE vocabulary table (BASELINE
    config 1: 16 learned concept embeddings vs a GPT-2-size table).  The generalisation of
    the pair loop that north_star names; returns (values [n,k], ids [n,k]) on the GPU."""
    dev = compute_device(embedding_matrix)
    table = to_kernel_dtype(embedding_matrix.detach()).to(dev)
    q = table[torch.as_tensor(list(query_token_ids), device=dev)]
    out = ops.concept_scan(q, table, k, inv_norm_t=inv_norm_table)
    return out.topk_val, out.topk_idx
