"""Tiny random-init stand-ins of the reference's models, shared by `oracle/gen_golden.py` (which
runs the REFERENCE's `MLLM.forward` / `evaluate_model` / vision loop arithmetic on them in the build
container and commits weights + outputs under tests/golden/) and by the GPU tests (which rebuild the
same architectures from `transformers`, load the golden weights, bind the drop-in forward and must
reproduce the golden loss / accuracy).  No `from_pretrained`: there is no network."""
from __future__ import annotations

import torch

LM_V, LM_H, N_VISION_TOKENS = 640, 64, 5


def tiny_lm_config():
    from transformers import Gemma3TextConfig
    return Gemma3TextConfig(vocab_size=LM_V, hidden_size=LM_H, intermediate_size=128,
                            num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                            head_dim=16, max_position_embeddings=128, sliding_window=64,
                            attn_implementation="eager")


def tiny_vit_config(num_labels=None):
    from transformers import ViTConfig
    kw = {} if num_labels is None else {"num_labels": num_labels}
    return ViTConfig(hidden_size=32, num_hidden_layers=1, num_attention_heads=2,
                     intermediate_size=64, image_size=32, patch_size=16, **kw)  # 4 patches + CLS = 5


class ToyTokenizer:
    """Deterministic word-piece tokenizer: splits on spaces and '-', maps each piece to an
    id by a fixed hash; unknown/empty text -> []. decode() joins 'yes'/'no'/tokN."""

    def __init__(self, vocab: int, yes_id: int = 7, no_id: int = 11):
        self.vocab, self.yes_id, self.no_id = vocab, yes_id, no_id

    def encode(self, text, add_special_tokens=False):
        pieces = [p for p in text.replace("-", " ").split(" ") if p]
        out = []
        for p in pieces:
            h = 0
            for ch in p:
                h = (h * 131 + ord(ch)) % 1000003
            out.append(h % self.vocab)
        return out

    def decode(self, ids, skip_special_tokens=True):
        words = {self.yes_id: "yes", self.no_id: "no"}
        return " ".join(words.get(int(i), f"tok{int(i)}") for i in ids)


class TinyMLLM(torch.nn.Module):
    """Same attribute names (hence the same state-dict keys) as the reference's `MLLM`
    (src/multimodal/mllm.py:14-88), built from configs instead of `from_pretrained`.  It has NO
    forward of its own: the tests bind `shims.mllm.fused_forward`, exactly the patch
    INTEGRATION.md documents for the reference class."""

    def __init__(self):
        super().__init__()
        from transformers import Gemma3ForCausalLM, ViTModel
        self.vision_model_name = "tiny-vit"
        self.language_model_name = "tiny-gemma3"
        self.num_vision_tokens = N_VISION_TOKENS
        self.vision_model = ViTModel(tiny_vit_config())
        self.language_model = Gemma3ForCausalLM(tiny_lm_config()).to(torch.bfloat16)
        self.projector = torch.nn.Linear(32, LM_H)
        self.tokenizer = ToyTokenizer(LM_V)
        self.labels_mapping = None
