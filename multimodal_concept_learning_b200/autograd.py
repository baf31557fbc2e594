"""Differentiable wrapper of the fused scan's cross-entropy: lets
``loss.backward()`` in the reference's training loops (``multimodal_training.py:140``,
``language_embed_only`` trains the table itself, ``mllm.py:181-184``; ``vision_training.py:120``)
run on top of the fused forward.

Forward: ``mcl_concept_scan`` (tcgen05 kernel, k = 1 epilogue) -- no ``[Q,V]`` logits, only
(m, s, sum_z, z_label) per row are kept.  Backward: ``mcl_ce_backward`` -- the library recomputes
the scores tile by tile on the tensor cores from the saved log-sum-exp, writes
d(loss)/dz = (softmax(z) - (1-eps) onehot - eps/V) * grad / n_valid in bf16 for an L2-sized block
at a time, and two hand-written tcgen05 GEMMs (MN-major operand descriptors, no transposed copies)
form d(loss)/dq = dz T and d(loss)/dT = dz^T q.  Only rows with a label take part (~1 % of the
positions under the reference's answer-only supervision).  PyTorch's part here is plumbing: the row
gather / scatter around the call and the casts of the fp32 gradients to the parameter dtype."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from ._lib import check, load

IGNORE_INDEX = -100


class FusedScanCrossEntropy(torch.autograd.Function):
    """mean_{valid rows} CE(softcap(scale * q table^T), labels) with label smoothing."""

    @staticmethod
    def forward(ctx, q, table, labels, scale: float, label_smoothing: float, softcap: float):
        out = ops.concept_scan(q.detach(), table.detach(), 1, normalize_q=False, normalize_t=False,
                               scale=scale, labels=labels, label_smoothing=label_smoothing,
                               softcap=softcap or None)
        ctx.save_for_backward(q, table, labels, out.lse)
        ctx.scale, ctx.eps, ctx.softcap = float(scale), float(label_smoothing), float(softcap or 0.0)
        ctx.mark_non_differentiable(out.topk_idx)
        return out.loss, out.topk_idx[:, 0]

    @staticmethod
    def backward(ctx, grad_loss, _grad_idx):
        q, table, labels, lse = ctx.saved_tensors
        need_q, need_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_q or need_t):
            return None, None, None, None, None, None
        dev = q.device
        rows = (labels != IGNORE_INDEX).nonzero().flatten()
        n = int(rows.numel())
        V, D = table.shape
        gq = torch.zeros(q.shape, dtype=torch.float32, device=dev) if need_q else None
        lib = load()
        # A bf16 table whose labelled rows fit one row block of the backward (<= 4096: the reference's
        # answer-only supervision labels ~1 % of the positions) gets its gradient in bf16 straight from
        # the last GEMM's epilogue: half the bytes of the step's largest write, no fp32 -> bf16 pass.
        gt_bf16 = bool(need_t and n > 0 and table.dtype == torch.bfloat16 and D % 8 == 0 and
                       n <= int(lib.mcl_ce_backward_block_rows(n, V, 0)))
        # (the library writes every element of the table gradient: no memset of [V, D] when n > 0)
        gt = (torch.empty if n > 0 else torch.zeros)((V, D), dtype=torch.bfloat16 if gt_bf16 else torch.float32,
                                                      device=dev) if need_t else None
        if n > 0:
            qv = ops._rowmajor(q.detach()[rows])              # the rows that carry a label
            tb = ops._rowmajor(table.detach())
            lse_v = lse[rows].contiguous()
            lab_v = labels[rows].contiguous()
            g = grad_loss.detach().to(torch.float32).reshape(1).contiguous()
            gq_v = torch.empty((n, D), dtype=torch.float32, device=dev) if need_q else None
            code = ops._dtype_code(qv)
            with torch.cuda.device(dev):
                ws_bytes = lib.mcl_ce_backward_workspace_bytes(n, V, D, code)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                check(lib.mcl_ce_backward_ex(qv.data_ptr(), tb.data_ptr(), code, n, V, D, qv.stride(0), tb.stride(0),
                                             lse_v.data_ptr(), lab_v.data_ptr(), ctx.scale, ctx.softcap, ctx.eps, V,
                                             g.data_ptr(), n, ops._ptr(gq_v), ops._ptr(gt), 0 if gt_bf16 else 1,
                                             ws.data_ptr(), ws_bytes, ops._stream(dev)))
            if need_q:
                gq[rows] = gq_v
        return (gq.to(q.dtype) if need_q else None, gt.to(table.dtype) if need_t else None,
                None, None, None, None)


def fused_cross_entropy(q: torch.Tensor, table: torch.Tensor, labels: torch.Tensor, *,
                        scale: float = 1.0, label_smoothing: float = 0.0, chunk_rows: Optional[int] = None,
                        softcap: Optional[float] = None):
    """Differentiable (w.r.t. ``q`` and ``table``) mean cross-entropy of ``scale * q table^T``
    against ``labels`` (``-100`` ignored); also returns the row-wise argmax.  ``q`` [Q,D] and
    ``table`` [V,D] are CUDA bf16/fp32 tensors of the same dtype.  ``softcap=c`` applies Gemma-2's
    ``c * tanh(z / c)`` to the logits.  (``chunk_rows`` is accepted for compatibility and ignored:
    the library blocks the backward itself.)"""
    labels = labels.to(device=q.device, dtype=torch.int64)
    return FusedScanCrossEntropy.apply(q, table, labels, scale, label_smoothing, float(softcap or 0.0))
