"""Differentiable wrapper of the fused scan's cross-entropy: lets
``loss.backward()`` in the reference's training loops (``multimodal_training.py:140``,
``language_embed_only`` trains the table itself, ``mllm.py:181-184``) run on top of the fused
forward.

Forward: ``mcl_concept_scan`` (tcgen05 kernel) -- no ``[Q,V]`` logits, only (m, s, sum_z, z_label)
per row are kept.  Backward: d(loss)/dz = (softmax(z) - (1-eps) onehot - eps/V) / n_valid on the
rows with a valid label only, recomputed chunk by chunk over the table from the saved
log-sum-exp, so memory stays O(rows_valid x chunk).  The three backward GEMMs are plain library
GEMMs (torch / cuBLAS): with the reference's answer-only supervision ~1 % of the rows carry a
label, which makes them skinny and HBM-bound; fusing them into a tcgen05 kernel is the next step
of this row, not done yet."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops

IGNORE_INDEX = -100


def _mm_f32(a: torch.Tensor, b_t: torch.Tensor) -> torch.Tensor:
    """a @ b_t.T with fp32 output (bf16 inputs keep fp32 accumulation visible to the softmax)."""
    try:
        return torch.mm(a, b_t.t(), out_dtype=torch.float32)
    except (TypeError, RuntimeError):
        return torch.mm(a.float(), b_t.float().t())


class FusedScanCrossEntropy(torch.autograd.Function):
    """mean_{valid rows} CE(scale * q @ table^T, labels) with label smoothing."""

    @staticmethod
    def forward(ctx, q, table, labels, scale: float, label_smoothing: float, chunk_rows: int):
        out = ops.concept_scan(q.detach(), table.detach(), 1, normalize_q=False, normalize_t=False,
                               scale=scale, labels=labels, label_smoothing=label_smoothing)
        valid = labels != IGNORE_INDEX
        ctx.save_for_backward(q, table, labels, out.lse, valid)
        ctx.scale, ctx.eps, ctx.chunk_rows = float(scale), float(label_smoothing), int(chunk_rows)
        ctx.mark_non_differentiable(out.topk_idx)
        return out.loss, out.topk_idx[:, 0]

    @staticmethod
    def backward(ctx, grad_loss, _grad_idx):
        q, table, labels, lse, valid = ctx.saved_tensors
        rows = valid.nonzero().flatten()
        n = rows.numel()
        need_q, need_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gq = torch.zeros_like(q, dtype=torch.float32) if need_q else None
        gt = torch.zeros_like(table, dtype=torch.float32) if need_t else None
        if n > 0 and (need_q or need_t):
            qv = q.detach()[rows]
            lse_v, lab_v = lse[rows], labels[rows]
            V = table.shape[0]
            coef = (grad_loss.float() / n) * ctx.scale          # d loss / d z, times dz/d(q.t)
            gqv = torch.zeros((n, q.shape[1]), dtype=torch.float32, device=q.device) if need_q else None
            for lo in range(0, V, ctx.chunk_rows):
                hi = min(V, lo + ctx.chunk_rows)
                tc = table.detach()[lo:hi]
                z = _mm_f32(qv, tc) * ctx.scale                  # [n, chunk] logits, fp32
                dz = torch.exp(z - lse_v[:, None])               # softmax via the saved LSE
                if ctx.eps:
                    dz -= ctx.eps / V
                inside = (lab_v >= lo) & (lab_v < hi)
                r_in = inside.nonzero().flatten()
                dz[r_in, lab_v[r_in] - lo] -= (1.0 - ctx.eps)
                dz *= coef
                dzl = dz.to(table.dtype)
                # skinny problems (the reference's answer-only supervision) stay in fp32; larger
                # ones use the table dtype for the two gradient GEMMs
                if need_q:
                    gqv += dz @ tc.float() if n <= 64 else torch.mm(dzl, tc).float()
                if need_t:
                    gt[lo:hi] = dz.t() @ qv.float() if n <= 64 else (dzl.t() @ qv).float()
            if need_q:
                gq[rows] = gqv
        return (gq.to(q.dtype) if need_q else None, gt.to(table.dtype) if need_t else None,
                None, None, None, None)


def fused_cross_entropy(q: torch.Tensor, table: torch.Tensor, labels: torch.Tensor, *,
                        scale: float = 1.0, label_smoothing: float = 0.0, chunk_rows: int = 32768,
                        softcap: Optional[float] = None):
    """Differentiable (w.r.t. ``q`` and ``table``) mean cross-entropy of ``scale * q @ table.T``
    against ``labels`` (``-100`` ignored); also returns the row-wise argmax.  ``q`` [Q,D] and
    ``table`` [V,D] are CUDA bf16/fp32 tensors of the same dtype."""
    labels = labels.to(device=q.device, dtype=torch.int64)
    if softcap:
        raise NotImplementedError("backward through soft-capped logits is not built yet")
    return FusedScanCrossEntropy.apply(q, table, labels, scale, label_smoothing, chunk_rows)
