"""Oracle: literal CPU restatements of the reference call sites on the hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Each function follows the cited
lines of ``/root/reference`` and is checked in ``tests/`` against (a) golden outputs
produced by the reference's own functions (``oracle/gen_golden.py``) and (b) the
batched formulation in ``concept_scan_ref`` that the CUDA path implements.
"""
from __future__ import annotations

import re
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from .concept_scan_ref import IGNORE_INDEX, row_inv_norm_ref


# -- a1: src/multimodal/token_embedding_analysis.py:175-181 ---------------------------
def extract_rgb_from_mapping_ref(labels_mapping: Dict[str, str], token: str):
    for rgb_key, token_value in labels_mapping.items():
        if token_value == token:
            hit = re.match(r"r(\d+)g(\d+)b(\d+)", rgb_key)
            if hit:
                r, g, b = map(int, hit.groups())
                return (r / 255.0, g / 255.0, b / 255.0)
    return (0.5, 0.5, 0.5)


def last_epoch_name_ref(embeddings_by_epoch) -> str:
    """token_embedding_analysis.py:200-208."""
    names = [n for n in embeddings_by_epoch if n.startswith("epoch_")]
    if not names:
        return "initial"
    return f"epoch_{max(int(n.split('_')[1]) for n in names)}"


def _cosine_pair_sklearn_semantics(a: np.ndarray, b: np.ndarray) -> np.float32:
    """What ``cosine_similarity([a],[b])[0][0]`` does for fp32 rows
    (sklearn/metrics/pairwise.py: normalize both, then dot; zero rows stay zero)."""
    tiny = np.float32(10.0) * np.finfo(np.float32).eps
    na = np.sqrt(np.dot(a, a), dtype=np.float32)
    nb = np.sqrt(np.dot(b, b), dtype=np.float32)
    na = np.float32(1.0) if na < tiny else na
    nb = np.float32(1.0) if nb < tiny else nb
    return np.float32(np.dot(a / na, b / nb))


def color_embedding_correlation_loop_ref(embeddings_by_epoch, ood_tokens, regular_tokens,
                                         ood_token_ids, regular_token_ids, labels_mapping):
    """token_embedding_analysis.py:183-260, restated pair by pair (the O(n^2) loop of
    :237-246) without sklearn, returning (r, color_distances, embedding_distances)."""
    table = embeddings_by_epoch[last_epoch_name_ref(embeddings_by_epoch)]
    ids = list(ood_token_ids) + list(regular_token_ids)
    names = list(ood_tokens) + list(regular_tokens)
    emb = table[ids].detach().cpu().float().numpy()                       # :220
    rgb = np.array([extract_rgb_from_mapping_ref(labels_mapping, t) for t in names])
    cd, ed = [], []
    for i in range(len(names)):
        for j in range(i + 1, len(names)):
            cd.append(np.sum(np.abs(rgb[i] - rgb[j])))                     # :240
            ed.append(1 - _cosine_pair_sklearn_semantics(emb[i], emb[j]))   # :244-245
    cd, ed = np.array(cd), np.array(ed)
    return np.corrcoef(cd, ed)[0, 1], cd, ed                              # :253


def pairwise_cosine_distance_sklearn_loop(token_embeddings: np.ndarray) -> np.ndarray:
    """The reference's LITERAL hot loop, token_embedding_analysis.py:237-246: one
    ``sklearn.metrics.pairwise.cosine_similarity([a], [b])`` call per pair i < j.  Returns the
    embedding distances in loop order.  (bench.py times it as the C1 CPU baseline.)"""
    from sklearn.metrics.pairwise import cosine_similarity
    n = len(token_embeddings)
    out = []
    for i in range(n):
        for j in range(i + 1, n):
            cos_sim = cosine_similarity([token_embeddings[i]], [token_embeddings[j]])[0][0]   # :244
            out.append(1 - cos_sim)                                                           # :245
    return np.array(out)


def color_embedding_correlation_batched_ref(embeddings_by_epoch, ood_tokens, regular_tokens,
                                            ood_token_ids, regular_token_ids, labels_mapping):
    """Same quantity from ONE n x n cosine matrix -- the formulation the CUDA shim uses
    (gather -> normalise -> E E^T -> upper triangle).  fp64 throughout."""
    table = embeddings_by_epoch[last_epoch_name_ref(embeddings_by_epoch)]
    ids = list(ood_token_ids) + list(regular_token_ids)
    names = list(ood_tokens) + list(regular_tokens)
    emb = table[ids].detach().cpu().double()
    e = emb * row_inv_norm_ref(table[ids].float(), torch.float64)[:, None]
    cos = (e @ e.T).numpy()
    rgb = np.array([extract_rgb_from_mapping_ref(labels_mapping, t) for t in names])
    iu = np.triu_indices(len(names), k=1)
    cd = np.abs(rgb[iu[0]] - rgb[iu[1]]).sum(axis=1)
    ed = 1.0 - cos[iu]
    return np.corrcoef(cd, ed)[0, 1], cd, ed


# -- a3: token_embedding_analysis_imagenet.py:261-286, multi_token.ipynb cell 2 --------
def average_embeddings_for_tokens_ref(tokenizer, embeddings_by_epoch, token_names
                                      ) -> Dict[str, torch.Tensor]:
    averaged: Dict[str, torch.Tensor] = {}
    if not embeddings_by_epoch:
        return averaged
    dim = next(iter(embeddings_by_epoch.values())).shape[1]
    for epoch_name, table in embeddings_by_epoch.items():
        if not token_names:
            averaged[epoch_name] = torch.empty((0, dim), dtype=table.dtype)
            continue
        rows = []
        for name in token_names:
            ids = tokenizer.encode(name, add_special_tokens=False)
            rows.append(table[ids].mean(dim=0) if ids else torch.zeros(dim, dtype=table.dtype))
        averaged[epoch_name] = torch.stack(rows)
    return averaged


def get_averaged_embedding_ref(text, tokenizer, embedding_matrix):
    """multi_token.ipynb cell 2 lines 1-12."""
    tokens = tokenizer.encode(text, add_special_tokens=False)
    return torch.mean(embedding_matrix[tokens], dim=0)


# -- a4/a6: mllm.py:115-120 -> HF lm_head + ForCausalLMLoss -----------------------------
def causal_lm_head_loss_ref(hidden: torch.Tensor, table: torch.Tensor, labels: torch.Tensor,
                            *, logits_dtype: Optional[torch.dtype] = None):
    """hidden [B,T,D] x tied table [V,D] -> (loss, logits [B,T,V]).
    modeling_gemma3.py:652 (``lm_head``), loss_utils.py:55-66 (upcast, pad+shift by one,
    flatten, ``F.cross_entropy(ignore_index=-100, reduction='mean')``).
    ``logits_dtype``: the dtype the LM head emits (bf16 under autocast in the reference;
    fp64/fp32 for a ground-truth run)."""
    dt = logits_dtype or hidden.dtype
    logits = (hidden.to(dt) @ table.to(dt).T)
    up = logits.float() if dt in (torch.bfloat16, torch.float16) else logits
    shift = F.pad(labels, (0, 1), value=IGNORE_INDEX)[..., 1:].contiguous()
    loss = F.cross_entropy(up.view(-1, table.shape[0]), shift.view(-1),
                           ignore_index=IGNORE_INDEX, reduction="mean")
    return loss, logits


# -- a5: multimodal_training.py:274-303 --------------------------------------------------
def evaluate_predictions_ref(logits: torch.Tensor, labels: torch.Tensor, tokenizer):
    """argmax over the vocab, UNSHIFTED mask ``labels != -100`` (the reference's quirk,
    :282), decode, yes/no string match.  Returns (correct, total, predicted_ids)."""
    predicted_ids = torch.argmax(logits, dim=-1)
    correct = total = 0
    for i in range(predicted_ids.size(0)):
        valid = labels[i] != IGNORE_INDEX
        if not valid.any():
            continue
        pred = predicted_ids[i][valid].cpu().tolist()
        true = labels[i][valid].cpu().tolist()
        pt = tokenizer.decode(pred, skip_special_tokens=True).strip()
        tt = tokenizer.decode(true, skip_special_tokens=True).strip()
        correct += int(("yes" in pt.lower()) == ("yes" in tt.lower()))
        total += 1
    return correct, total, predicted_ids


# -- a7: vision_training.py:80-83,116,132 -------------------------------------------------
def vision_ce_top1_ref(features: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                       labels: torch.Tensor, label_smoothing: float = 0.0):
    """logits = Linear(features); CrossEntropyLoss(label_smoothing) mean; torch.max(.,1);
    returns (loss, predicted, n_correct)."""
    logits = F.linear(features.float(), weight.float(), None if bias is None else bias.float())
    crit = torch.nn.CrossEntropyLoss(label_smoothing=label_smoothing) if label_smoothing > 0 \
        else torch.nn.CrossEntropyLoss()
    loss = crit(logits, labels)
    _, predicted = torch.max(logits.data, 1)
    return loss, predicted, int((predicted == labels).sum())
