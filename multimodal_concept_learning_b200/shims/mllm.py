"""Shim for the LM head + loss that ``MLLM.forward`` reaches through
``self.language_model(..., labels=labels)`` (``src/multimodal/mllm.py:115-120`` ->
``modeling_gemma3.py:649-664`` -> ``loss_utils.py:45-67``).

The reference materialises ``[B,T,V]`` bf16 logits, upcasts them to fp32 and runs
``F.cross_entropy``; here hidden states go straight into the fused scan: loss from the
online log-sum-exp, predictions from the k=1 top-k, logits never stored.  The LM head is
the input-embedding table (tied, ``modeling_gemma3.py:593``): raw dot product, no
normalisation, scale 1."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn.functional as F

from .. import ops
from ._common import compute_device, to_kernel_dtype
from .lazy_logits import LazyLogits

IGNORE_INDEX = -100


@dataclass
class FusedCausalLMOutput:
    """The fields of HF's ``CausalLMOutputWithPast`` the reference's loops read: ``.loss``
    (``multimodal_training.py:137,169,271``) and ``.logits`` (``:274``).  ``logits`` is a
    :class:`LazyLogits`: ``torch.argmax(outputs.logits, dim=-1)`` is answered by the fused k = 1
    scan, any other use materialises the real ``[B,T,V]`` tensor on demand."""
    loss: Optional[torch.Tensor]
    predicted_ids: torch.Tensor          # argmax over the vocabulary, [B,T] (-1 where not computed)
    logits: Optional[LazyLogits] = None

    def __getitem__(self, key):          # ModelOutput-style access: outputs["loss"], outputs[0]
        if isinstance(key, str):
            return getattr(self, key)
        return tuple(v for v in (self.loss, self.logits) if v is not None)[key]


def shift_labels(labels: torch.Tensor, ignore_index: int = IGNORE_INDEX) -> torch.Tensor:
    """loss_utils.py:57-59: pad one ignore on the right, drop the first -> token t predicts t+1."""
    return F.pad(labels, (0, 1), value=ignore_index)[..., 1:].contiguous()


def lm_head_loss_and_argmax(hidden_states: torch.Tensor, embedding_table: torch.Tensor,
                            labels: Optional[torch.Tensor] = None, *, rows: str = "labelled",
                            inv_norm_table=None, softcap: Optional[float] = None) -> FusedCausalLMOutput:
    """hidden [B,T,D] x table [V,D].  ``rows``: "all" scans every position (what HF computes);
    "labelled" scans only positions that the loss (shifted mask) or the reference's accuracy
    (unshifted mask, multimodal_training.py:282) reads -- identical loss and accuracy, ~100x
    fewer query rows with the reference's answer-only supervision."""
    dev = compute_device(hidden_states, embedding_table)
    B, T, D = hidden_states.shape
    # training (run_training, multimodal_training.py:131-140): keep the autograd graph and route
    # the loss through the differentiable wrapper; evaluation: plain fused scan
    train = torch.is_grad_enabled() and labels is not None and (
        hidden_states.requires_grad or embedding_table.requires_grad)
    _d = (lambda t: t) if train else (lambda t: t.detach())
    h = to_kernel_dtype(_d(hidden_states)).to(dev).reshape(B * T, D)
    table = to_kernel_dtype(_d(embedding_table)).to(dev)
    if h.dtype != table.dtype:
        h = h.to(table.dtype)
    shifted = None
    if labels is not None:
        labels = labels.to(dev)
        shifted = shift_labels(labels).reshape(-1)
    if rows == "all" or labels is None:
        sel = None
        q, lab = h, shifted
    else:
        need = (shifted != IGNORE_INDEX) | (labels.reshape(-1) != IGNORE_INDEX)
        sel = need.nonzero().flatten()
        q, lab = h[sel], shifted[sel]
    pred = torch.full((B * T,), -1, dtype=torch.int64, device=dev)
    loss = None
    if q.shape[0] > 0 and train:
        from ..autograd import fused_cross_entropy
        loss, top1 = fused_cross_entropy(q, table, lab, softcap=softcap)
        if sel is None:
            pred = top1
        else:
            pred[sel] = top1
    elif q.shape[0] > 0:
        out = ops.concept_scan(q, table, 1, normalize_q=False, normalize_t=False, labels=lab, softcap=softcap)
        if sel is None:
            pred = out.topk_idx[:, 0]
        else:
            pred[sel] = out.topk_idx[:, 0]
        if lab is not None:
            loss = out.loss           # mean over rows with a shifted label (reduction='mean')
    elif labels is not None:
        loss = torch.tensor(float("nan"), device=dev)    # F.cross_entropy over zero valid rows
    return FusedCausalLMOutput(loss=loss, predicted_ids=pred.reshape(B, T),
                               logits=LazyLogits(hidden_states, embedding_table, softcap=softcap))


def fused_forward(self, images, input_ids, attention_mask, labels=None):
    """Drop-in for ``MLLM.forward`` (``src/multimodal/mllm.py:90-121``): same arguments, an output
    with the ``.loss`` / ``.logits`` the training and evaluation loops read.  Bind it with

        from multimodal_concept_learning_b200.shims.mllm import fused_forward
        MLLM.forward = fused_forward

    Lines :92-112 (vision encoder -> projector -> token embeddings with the first
    ``num_vision_tokens`` positions overwritten) are the reference's own statements.  Line :115
    ``self.language_model(inputs_embeds, attention_mask, labels)`` is split: the decoder stack
    ``self.language_model.model`` runs as before, and the tied LM head + fp32 upcast +
    ``F.cross_entropy`` + (later) ``argmax`` that HF would run on ``[B,T,V]`` logits
    (``modeling_gemma3.py:649-664``, ``loss_utils.py:45-67``) become ONE fused scan over the
    positions that carry a label -- loss from the online log-sum-exp, predictions from the k = 1
    epilogue.  Under ``torch.enable_grad()`` the loss is differentiable with respect to the hidden
    states and the table (``language_embed_only`` trains the table, :181-184)."""
    if "timm" in self.vision_model_name:
        image_embeds = self.vision_model.timm_model.forward_features(images)
    else:
        vision_outputs = self.vision_model(pixel_values=images, output_hidden_states=True, return_dict=True)
        image_embeds = vision_outputs.last_hidden_state
    projected_image_embeds = self.projector(image_embeds)
    input_embedding_layer = self.language_model.get_input_embeddings()
    language_embeds = input_embedding_layer(input_ids)
    language_embeds[:, :self.num_vision_tokens, :] = projected_image_embeds
    decoder = self.language_model.model
    hidden = decoder(inputs_embeds=language_embeds, attention_mask=attention_mask, return_dict=True).last_hidden_state
    head = self.language_model.get_output_embeddings()          # lm_head, tied to the input table
    table = head.weight if head is not None else input_embedding_layer.weight
    cfg = self.language_model.config
    softcap = getattr(cfg, "final_logit_softcapping", None) or getattr(getattr(cfg, "text_config", None),
                                                                       "final_logit_softcapping", None)
    return lm_head_loss_and_argmax(hidden, table, labels, rows="labelled", softcap=softcap)

