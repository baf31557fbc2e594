"""Oracle: the fused concept scan, restated with the primitives the reference calls.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference never names a "similarity scan"; SURVEY.md section 8a maps the path to
four call sites.  This module composes them into one function with the exact
semantics the CUDA path must reproduce:

* row L2 normalisation with scikit-learn semantics -- rows whose norm is below
  ``10 * eps`` are divided by 1 (``sklearn/preprocessing/_data.py`` ``normalize`` ->
  ``_handle_zeros_in_scale``), the call site being
  ``/root/reference/src/multimodal/token_embedding_analysis.py:244``;
* ``scores = q @ table.T * scale`` -- the tied LM head of
  ``/root/reference/src/multimodal/mllm.py:115`` (raw dot product, scale 1);
* row-wise ``torch.topk`` (k=1 is the ``torch.argmax`` of
  ``/root/reference/src/multimodal/multimodal_training.py:276`` and the
  ``torch.max(..., 1)`` of ``/root/reference/src/vision/vision_training.py:132``);
* ``F.cross_entropy(ignore_index=-100, label_smoothing=eps)`` -- HF
  ``ForCausalLMLoss`` behind ``mllm.py:115`` and ``nn.CrossEntropyLoss`` at
  ``vision_training.py:81-83``.

Ties: exact score ties are real in the reference (``mllm.py:73`` copies rows), and
``torch.topk`` does not define which duplicate wins.  The contract here, and in the
CUDA kernels, is *lowest index wins*; comparators in ``tests/`` are tie-agnostic
where they compare against ``torch.topk`` itself.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

IGNORE_INDEX = -100
_TINY = 10.0 * float(np.finfo(np.float32).eps)  # sklearn's zero-norm guard for fp32


def row_inv_norm_ref(x: torch.Tensor, dtype: torch.dtype = torch.float64) -> torch.Tensor:
    """1/||row||_2 with sklearn's zero handling (norm < 10*eps -> 1)."""
    xf = x.detach().to("cpu").to(dtype)
    n = torch.sqrt((xf * xf).sum(dim=1))
    n = torch.where(n < _TINY, torch.ones_like(n), n)
    return 1.0 / n


@dataclass
class ScanResult:
    topk_val: torch.Tensor          # [Q,k] descending
    topk_idx: torch.Tensor          # [Q,k] int64, global row ids
    m: torch.Tensor                 # [Q] row max of z
    s: torch.Tensor                 # [Q] sum exp(z - m)
    sum_z: torch.Tensor             # [Q]
    z_label: torch.Tensor           # [Q] (0 where label is ignored / not local)
    lse: torch.Tensor               # [Q]
    loss_rows: Optional[torch.Tensor] = None   # [Q], 0 on ignored rows
    loss: Optional[torch.Tensor] = None        # scalar, mean over valid rows
    scores: Optional[torch.Tensor] = None      # [Q,V] (kept only if asked)


def scores_ref(q, table, *, normalize_q=True, normalize_t=True, scale=1.0,
               dtype=torch.float64, softcap=None) -> torch.Tensor:
    qf = q.detach().to("cpu").to(dtype)
    tf = table.detach().to("cpu").to(dtype)
    z = qf @ tf.T
    if normalize_q:
        z = z * row_inv_norm_ref(q, dtype)[:, None]
    if normalize_t:
        z = z * row_inv_norm_ref(table, dtype)[None, :]
    z = z * scale
    if softcap:   # modeling_gemma3.py:653-656: logits / c -> tanh -> * c
        z = torch.tanh(z / softcap) * softcap
    return z


def topk_lowest_index(z: torch.Tensor, k: int):
    """Row-wise top-k, descending, exact ties broken by lowest index."""
    # stable sort on descending values keeps the lower index first among equals
    order = torch.sort(z, dim=1, descending=True, stable=True).indices[:, :k]
    return torch.gather(z, 1, order), order


def stats_from_scores(z: torch.Tensor, labels: Optional[torch.Tensor], index_base: int = 0):
    m = z.max(dim=1).values
    s = torch.exp(z - m[:, None]).sum(dim=1)
    sum_z = z.sum(dim=1)
    z_label = torch.zeros_like(m)
    if labels is not None:
        local = labels.to(torch.int64) - index_base
        ok = (labels != IGNORE_INDEX) & (local >= 0) & (local < z.shape[1])
        rows = torch.nonzero(ok).flatten()
        z_label[rows] = z[rows, local[rows]]
    return m, s, sum_z, z_label


def loss_from_stats(lse, sum_z, z_label, labels, vocab: int, label_smoothing: float = 0.0):
    """CE with label smoothing from per-row stats:
    (1-eps)*(lse - z_y) + eps*(lse - sum_z/V); ignored rows contribute 0; mean over valid."""
    valid = labels != IGNORE_INDEX
    rows = (1.0 - label_smoothing) * (lse - z_label) + label_smoothing * (lse - sum_z / vocab)
    rows = torch.where(valid, rows, torch.zeros_like(rows))
    n = int(valid.sum())
    loss = rows.sum() / n if n > 0 else torch.tensor(float("nan"), dtype=rows.dtype)
    return rows, loss


def concept_scan_ref(q, table, k: int, *, normalize_q=True, normalize_t=True, scale=1.0,
                     labels: Optional[torch.Tensor] = None, label_smoothing: float = 0.0,
                     index_base: int = 0, vocab_total: Optional[int] = None,
                     dtype=torch.float64, keep_scores=False, softcap=None) -> ScanResult:
    """The whole path on CPU.  ``dtype=float64`` is the ground truth on the given
    (possibly bf16) input values; ``float32`` is what the reference's own fp32
    composition computes."""
    z = scores_ref(q, table, normalize_q=normalize_q, normalize_t=normalize_t, scale=scale,
                   dtype=dtype, softcap=softcap)
    V = z.shape[1]
    if not 1 <= k <= V:
        raise ValueError(f"k={k} out of range for V={V}")
    val, idx = topk_lowest_index(z, k)
    m, s, sum_z, z_label = stats_from_scores(z, labels, index_base)
    lse = m + torch.log(s)
    res = ScanResult(val, idx + index_base, m, s, sum_z, z_label, lse)
    if labels is not None:
        res.loss_rows, res.loss = loss_from_stats(
            lse, sum_z, z_label, labels, vocab_total or V, label_smoothing)
    if keep_scores:
        res.scores = z
    return res


def prepare_table_ref(table, normalize=True):
    """Per-table-version work of the composition below (fp32 copy + row normalisation), so that a
    timing of the per-query-batch step can hoist it the way the GPU path caches 1/||row||."""
    tf = table.detach().to("cpu").float()
    return F.normalize(tf, dim=1) if normalize else tf


def torch_composition_ref(q, table, k, *, normalize=True, scale=1.0, labels=None,
                          label_smoothing=0.0, table_prepared=False):
    """The literal PyTorch composition the reference's training loops reduce to
    (F.normalize / @ / topk / cross_entropy), fp32 on CPU.  Used as cpu_baseline and to
    check ``concept_scan_ref`` against library kernels rather than against itself.
    ``table_prepared``: ``table`` is the output of :func:`prepare_table_ref`."""
    qf = q.detach().to("cpu").float()
    if normalize:
        qf = F.normalize(qf, dim=1)
    tf = table if table_prepared else prepare_table_ref(table, normalize)
    z = (qf @ tf.T) * scale
    val, idx = torch.topk(z, k, dim=1)
    out = {"topk_val": val, "topk_idx": idx, "lse": torch.logsumexp(z, dim=1)}
    if labels is not None:
        out["loss"] = F.cross_entropy(z, labels, ignore_index=IGNORE_INDEX,
                                      label_smoothing=label_smoothing)
    return out


# ----------------------------------------------------------------------------------
# vocab-row sharding and the merge
# ----------------------------------------------------------------------------------

def shard_bounds(V: int, world: int) -> List[tuple]:
    per = -(-V // world)
    return [(min(V, r * per), min(V, (r + 1) * per)) for r in range(world)]


def merge_ref(vals: Sequence[torch.Tensor], idxs: Sequence[torch.Tensor],
              ms, ss, sum_zs, z_labels, k: int):
    """Merge per-shard partial results.  Candidates: concatenate, order by
    (value desc, index asc), keep k.  LSE: m* = max m_r, s* = sum s_r exp(m_r - m*)."""
    val = torch.cat(list(vals), dim=1)
    idx = torch.cat(list(idxs), dim=1)
    # order by (value desc, index asc): sort by index first, then stable by value
    o1 = torch.sort(idx, dim=1, stable=True).indices
    val1, idx1 = torch.gather(val, 1, o1), torch.gather(idx, 1, o1)
    o2 = torch.sort(val1, dim=1, descending=True, stable=True).indices[:, :k]
    out_val, out_idx = torch.gather(val1, 1, o2), torch.gather(idx1, 1, o2)
    M = torch.stack(list(ms), 0)
    m = M.max(dim=0).values
    s = (torch.stack(list(ss), 0) * torch.exp(M - m[None, :])).sum(0)
    sum_z = torch.stack(list(sum_zs), 0).sum(0)
    z_label = torch.stack(list(z_labels), 0).sum(0)
    return out_val, out_idx, m, s, sum_z, z_label


def concept_scan_sharded_ref(q, table, k, world: int, **kw) -> ScanResult:
    """Single-process simulation of the N-rank vocab-sharded scan + merge."""
    labels = kw.get("labels")
    ls = kw.pop("label_smoothing", 0.0)
    V = table.shape[0]
    parts = []
    for lo, hi in shard_bounds(V, world):
        if hi <= lo:
            continue
        kk = min(k, hi - lo)
        r = concept_scan_ref(q, table[lo:hi], kk, index_base=lo, vocab_total=V, **kw)
        if kk < k:  # pad short shards with -inf candidates
            pad = k - kk
            r.topk_val = torch.cat([r.topk_val, torch.full((q.shape[0], pad), -math.inf,
                                                           dtype=r.topk_val.dtype)], 1)
            r.topk_idx = torch.cat([r.topk_idx, torch.full((q.shape[0], pad), -1,
                                                           dtype=torch.int64)], 1)
        parts.append(r)
    val, idx, m, s, sum_z, z_label = merge_ref(
        [p.topk_val for p in parts], [p.topk_idx for p in parts], [p.m for p in parts],
        [p.s for p in parts], [p.sum_z for p in parts], [p.z_label for p in parts], k)
    lse = m + torch.log(s)
    res = ScanResult(val, idx, m, s, sum_z, z_label, lse)
    if labels is not None:
        res.loss_rows, res.loss = loss_from_stats(lse, sum_z, z_label, labels, V, ls)
    return res


# ----------------------------------------------------------------------------------
# multi-token concept embeddings: gather rows + mean (+ L2 normalise)
# ----------------------------------------------------------------------------------

def gather_mean_ref(table: torch.Tensor, offsets, ids, *, normalize=False,
                    out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """CSR form of ``average_embeddings_for_tokens``
    (``/root/reference/src/multimodal/token_embedding_analysis_imagenet.py:277-284``):
    row i = mean(table[ids[offsets[i]:offsets[i+1]]]); empty -> zeros (``:283``).
    Accumulates in fp32, rounds once to ``out_dtype`` (default: the table dtype, as the
    reference's ``.mean(dim=0)`` on a bf16 tensor does)."""
    offsets = [int(o) for o in offsets]
    ids_t = torch.as_tensor(ids, dtype=torch.int64)
    out_dtype = out_dtype or table.dtype
    rows = []
    for i in range(len(offsets) - 1):
        sel = ids_t[offsets[i]:offsets[i + 1]]
        if sel.numel() == 0:
            rows.append(torch.zeros(table.shape[1], dtype=torch.float32))
        else:
            rows.append(table[sel].float().sum(0) / sel.numel())
    out = torch.stack(rows) if rows else torch.zeros((0, table.shape[1]))
    if normalize:
        # multi_token.ipynb cell 3 line 16: x / ||x|| ; zero rows guarded as sklearn does
        out = out * row_inv_norm_ref(out, torch.float32)[:, None]
    return out.to(out_dtype)
