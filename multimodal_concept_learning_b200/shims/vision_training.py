"""Shim for the classifier-head arithmetic of ``src/vision/vision_training.py``:
``criterion(outputs.logits, labels)`` (:81-83,116) and ``torch.max(outputs.logits.data, 1)``
(:132,153,228).  The bias of ``Linear(768, C)`` rides along as one extra K column."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .. import ops
from ._common import compute_device, to_kernel_dtype


def classifier_loss_and_top1(features: torch.Tensor, weight: torch.Tensor,
                             bias: Optional[torch.Tensor], labels: torch.Tensor,
                             label_smoothing: float = 0.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """features [B,D] x weight [C,D] (+ bias [C]) -> (mean CE with label smoothing, argmax [B]).
    First-max-wins ties, as ``torch.max`` does."""
    dev = compute_device(features, weight)
    f = to_kernel_dtype(features.detach()).to(dev)
    w = to_kernel_dtype(weight.detach()).to(device=dev, dtype=f.dtype)
    if bias is not None:
        pad = 8 if f.dtype == torch.bfloat16 else 4          # keep rows 16-byte aligned
        fe = torch.zeros((f.shape[0], pad), dtype=f.dtype, device=dev)
        fe[:, 0] = 1
        we = torch.zeros((w.shape[0], pad), dtype=w.dtype, device=dev)
        we[:, 0] = bias.detach().to(device=dev, dtype=w.dtype)
        f, w = torch.cat([f, fe], 1), torch.cat([w, we], 1)
    out = ops.concept_scan(f, w, 1, normalize_q=False, normalize_t=False, labels=labels.to(dev),
                           label_smoothing=label_smoothing)
    return out.loss, out.topk_idx[:, 0]
