"""Shim for the classifier-head arithmetic of ``src/vision/vision_training.py``:
``criterion(outputs.logits, labels)`` (:81-83,116) and ``torch.max(outputs.logits.data, 1)``
(:132,153,228).  The bias of ``Linear(768, C)`` rides along as one extra K column."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from dataclasses import dataclass

from .. import ops
from ._common import compute_device, to_kernel_dtype
from .lazy_logits import LazyLogits


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


class FusedCrossEntropyAndTop1(torch.nn.Module):
    """Drop-in for the ``criterion`` of the vision loops (``vision_training.py:80-83,213-215``):
    ``criterion = FusedCrossEntropyAndTop1(label_smoothing=config.label_smoothing)``, then
    ``criterion(outputs.logits, labels)`` (:116,149,223) as before.  With
    ``ViTForImageClassification.forward = fused_vit_forward`` bound, ``outputs.logits`` is a
    :class:`LazyLogits` and the loss comes from the fused scan (differentiable in training);
    ``torch.max(outputs.logits.data, 1)`` (:132,153,228) is answered by the same object.  Dense
    logits are rejected loudly: this module is the fused path, not a wrapper around torch's."""

    def __init__(self, label_smoothing: float = 0.0):
        super().__init__()
        self.label_smoothing = float(label_smoothing)

    def forward(self, logits, labels):
        from .lazy_logits import LazyLogits
        if not isinstance(logits, LazyLogits):
            raise TypeError("FusedCrossEntropyAndTop1 takes the LazyLogits of a fused forward "
                            "(bind shims.vision_training.fused_vit_forward); it has no dense fallback")
        return logits.cross_entropy(labels, label_smoothing=self.label_smoothing)


@dataclass
class FusedImageClassifierOutput:
    """``ImageClassifierOutput`` fields the reference's loops read (``.logits``, ``.loss``)."""
    logits: object
    loss: Optional[torch.Tensor] = None


def fused_vit_forward(self, pixel_values=None, labels=None, **kwargs):
    """Drop-in for ``ViTForImageClassification.forward`` as the reference calls it
    (``outputs = model(images)``, ``vision_training.py:115,148,222``): the encoder ``self.vit`` runs
    as before, the ``Linear(hidden, num_labels)`` head on the CLS token (``modeling_vit.py:642``) is
    NOT applied -- ``outputs.logits`` is a :class:`LazyLogits` over (CLS features, classifier weight,
    bias) that the criterion and ``torch.max`` consume fused.  fp16 / fp32 features (the launch uses
    fp16 autocast, ``scripts/train_vision_accelerate.sh:43``) take the fp32 check-path kernel,
    exactly (fp16 -> fp32 is exact); bf16 features take the tcgen05 path."""
    outputs = self.vit(pixel_values, **kwargs)
    cls = outputs.last_hidden_state[:, 0, :]
    lazy = LazyLogits(cls, self.classifier.weight, self.classifier.bias)
    loss = None
    if labels is not None:
        loss = lazy.cross_entropy(labels)
    return FusedImageClassifierOutput(logits=lazy, loss=loss)
