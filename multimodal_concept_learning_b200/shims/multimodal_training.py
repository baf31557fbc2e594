"""Shim for the evaluation arithmetic of ``src/multimodal/multimodal_training.py:250-316``."""
from __future__ import annotations

import torch

from .mllm import IGNORE_INDEX, lm_head_loss_and_argmax


def count_yes_no_matches(predicted_ids: torch.Tensor, labels: torch.Tensor, tokenizer):
    """Reference :278-303 verbatim in meaning: per sample keep positions where the UNSHIFTED
    labels != -100, decode prediction and truth, compare "yes" membership."""
    correct = total = 0
    predicted_ids, labels = predicted_ids.cpu(), labels.cpu()
    for i in range(predicted_ids.size(0)):
        valid_mask = labels[i] != IGNORE_INDEX
        if not valid_mask.any():
            continue
        pred_tokens = predicted_ids[i][valid_mask].tolist()
        true_tokens = labels[i][valid_mask].tolist()
        pred_text = tokenizer.decode(pred_tokens, skip_special_tokens=True).strip()
        true_text = tokenizer.decode(true_tokens, skip_special_tokens=True).strip()
        if ("yes" in pred_text.lower()) == ("yes" in true_text.lower()):
            correct += 1
        total += 1
    return correct, total


def evaluate_hidden_batches(batches, embedding_table: torch.Tensor, tokenizer):
    """``evaluate_model`` with the LM head fused: ``batches`` yields dicts with
    ``hidden_states`` [B,T,D] (the decoder's last hidden state) and ``labels`` [B,T].
    Returns the reference's ``{'test_loss', 'test_acc'}`` (:313-316)."""
    test_loss, correct, total, n = 0.0, 0, 0, 0
    for batch in batches:
        out = lm_head_loss_and_argmax(batch["hidden_states"], embedding_table, batch["labels"])
        test_loss += out.loss.item()
        c, t = count_yes_no_matches(out.predicted_ids, batch["labels"], tokenizer)
        correct, total, n = correct + c, total + t, n + 1
    return {"test_loss": test_loss / max(n, 1),
            "test_acc": 100.0 * correct / total if total > 0 else 0.0}


def evaluate_model(model, test_loader, config, accelerator):
    """Drop-in for ``evaluate_model(model, test_loader, config, accelerator)``
    (``src/multimodal/multimodal_training.py:250-316``): same arguments, same prints, same
    ``{'test_loss', 'test_acc'}``.  With ``MLLM.forward = shims.mllm.fused_forward`` bound, the
    outputs carry ``predicted_ids`` (the fused argmax at the positions the accuracy reads, i.e. the
    UNSHIFTED label mask of :282); a model with the stock forward still works through
    ``torch.argmax(outputs.logits, dim=-1)`` (:276)."""
    model.eval()
    test_loss = 0.0
    correct_predictions = 0
    total_predictions = 0
    unwrapped_model = accelerator.unwrap_model(model)
    with torch.no_grad():
        for batch in test_loader:
            with accelerator.autocast():
                outputs = model(images=batch["images"], input_ids=batch["input_ids"],
                                attention_mask=batch["attention_mask"], labels=batch["labels"])
                test_loss += outputs.loss.item()
                labels = batch["labels"]
                predicted_ids = getattr(outputs, "predicted_ids", None)
                if predicted_ids is None:
                    predicted_ids = torch.argmax(outputs.logits, dim=-1)
                c, t = count_yes_no_matches(predicted_ids, labels, unwrapped_model.tokenizer)
                correct_predictions += c
                total_predictions += t
    test_loss /= len(test_loader)
    test_acc = 100. * correct_predictions / total_predictions if total_predictions > 0 else 0.0
    if accelerator.is_main_process:
        print(f"Test Results:")
        print(f"Test Loss: {test_loss:.4f}")
        print(f"Test Accuracy: {test_acc:.2f}%")
    return {'test_loss': test_loss, 'test_acc': test_acc}
