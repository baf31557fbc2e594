"""CPU: the host-side logic of the drop-in shims (everything around the kernels) against the
reference's statements as restated in oracle/reference_sites.py and against the third-party
functions the reference calls (HF ``ForCausalLMLoss``).  No kernel runs here; on a machine
without a GPU the compute entry points must refuse loudly instead of computing on the CPU."""
import json
import os

import pytest
import torch

from multimodal_concept_learning_b200.shims import mllm as shim_mllm
from multimodal_concept_learning_b200.shims import multimodal_training as shim_mt
from multimodal_concept_learning_b200.shims import token_embedding_analysis as shim_tea
from multimodal_concept_learning_b200.shims import token_embedding_analysis_imagenet as shim_img
from oracle import reference_sites as R
from tests.tiny_models import ToyTokenizer

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NO_GPU = not torch.cuda.is_available()


def test_shift_labels_is_what_hf_causal_lm_loss_does():
    """loss_utils.py:57-59: the loss of position t is taken against the label of t+1.  The shim's
    shifted labels + a plain cross-entropy must equal HF's own ForCausalLMLoss on the same logits."""
    from transformers.loss.loss_utils import ForCausalLMLoss
    g = torch.Generator().manual_seed(3)
    B, T, V = 3, 17, 29
    logits = torch.randn(B, T, V, generator=g)
    labels = torch.randint(0, V, (B, T), generator=g)
    labels[:, :9] = -100                                    # prompt positions carry no label
    labels[1, 12] = -100
    shifted = shim_mllm.shift_labels(labels)
    assert shifted.shape == labels.shape
    assert torch.equal(shifted[:, :-1], labels[:, 1:]) and bool((shifted[:, -1] == -100).all())
    mine = torch.nn.functional.cross_entropy(logits.reshape(-1, V), shifted.reshape(-1), ignore_index=-100)
    want = ForCausalLMLoss(logits, labels, vocab_size=V)
    torch.testing.assert_close(mine, want, rtol=1e-6, atol=1e-6)
    # and the oracle's restatement of the same lines agrees (it is what the GPU tests compare with)
    hidden = torch.randn(B, T, 8, generator=g)
    table = torch.randn(V, 8, generator=g)
    loss_ref, lg = R.causal_lm_head_loss_ref(hidden, table, labels, logits_dtype=torch.float32)
    torch.testing.assert_close(loss_ref, ForCausalLMLoss(lg, labels, vocab_size=V), rtol=1e-6, atol=1e-6)


def test_yes_no_accuracy_keeps_the_unshifted_mask_quirk():
    """multimodal_training.py:276-303: predictions at the positions where the UNSHIFTED labels are
    set are decoded and compared by "yes" membership.  Shim (takes predicted ids) vs the oracle's
    restatement (takes logits), on cases that a shifted mask would score differently."""
    tok = ToyTokenizer(50)
    g = torch.Generator().manual_seed(11)
    B, T, V = 6, 12, 50
    labels = torch.full((B, T), -100, dtype=torch.int64)
    labels[0, 9:11] = torch.tensor([tok.yes_id, 3])
    labels[1, 9:11] = torch.tensor([tok.no_id, 3])
    labels[2, 10] = tok.yes_id
    labels[4, 8:11] = torch.tensor([5, tok.no_id, 3])       # row 3 and 5 carry no label at all
    logits = torch.randn(B, T, V, generator=g)
    logits[0, 9, tok.yes_id] = 50.0                         # right at the unshifted position
    logits[1, 8, tok.no_id] = 50.0                          # right only under a SHIFTED reading
    logits[1, 9, tok.yes_id] = 50.0                         # ... and wrong where the reference looks
    logits[2, 10, tok.no_id] = 50.0
    logits[4, 9, tok.no_id] = 50.0
    c_ref, t_ref, pred = R.evaluate_predictions_ref(logits, labels, tok)
    c, t = shim_mt.count_yes_no_matches(pred, labels, tok)
    assert (c, t) == (c_ref, t_ref) == (2, 4)
    # predictions at unlabelled positions never matter (the fused path leaves them at -1)
    sparse = torch.where(labels != -100, pred, torch.full_like(pred, -1))
    assert shim_mt.count_yes_no_matches(sparse, labels, tok) == (c_ref, t_ref)


def test_tokens_to_csr_matches_per_name_encoding():
    tok = ToyTokenizer(1000)
    names = ["red", "dark-olive green", "", "sky blue", "-", "r255g32b0"]
    offsets, ids = shim_img.tokens_to_csr(tok, names)
    assert offsets.dtype == ids.dtype == torch.int64 and offsets.numel() == len(names) + 1
    assert int(offsets[0]) == 0 and int(offsets[-1]) == ids.numel()
    for i, name in enumerate(names):
        assert ids[int(offsets[i]):int(offsets[i + 1])].tolist() == tok.encode(name, add_special_tokens=False)
    assert int(offsets[3]) == int(offsets[2])               # the empty name owns no ids (reference: zeros row)


def test_average_embeddings_empty_inputs_need_no_device():
    """token_embedding_analysis_imagenet.py:261-286 on empty inputs: {} for no epochs, [0, D] per
    epoch for no names -- decided on the host, as in the reference."""
    tok = ToyTokenizer(100)
    assert shim_img.average_embeddings_for_tokens(tok, {}, ["red"]) == {}
    tables = {"initial": torch.zeros(100, 16, dtype=torch.bfloat16), "epoch_0": torch.ones(100, 16, dtype=torch.bfloat16)}
    out = shim_img.average_embeddings_for_tokens(tok, tables, [])
    assert list(out) == ["initial", "epoch_0"]
    assert all(v.shape == (0, 16) and v.dtype == torch.bfloat16 for v in out.values())
    want = R.average_embeddings_for_tokens_ref(tok, tables, [])
    assert {k: tuple(v.shape) for k, v in want.items()} == {k: tuple(v.shape) for k, v in out.items()}


def test_extract_rgb_matches_the_reference_parser_on_the_golden_mapping():
    mapping = json.load(open(os.path.join(GOLDEN, "12_colors_3k_labels_mapping.json")))
    tokens = sorted(set(mapping.values())) + ["<not a token>"]
    for tkn in tokens:
        assert shim_tea.extract_rgb_from_mapping(mapping, tkn) == R.extract_rgb_from_mapping_ref(mapping, tkn)


def test_causal_lm_output_indexing_like_hf_model_output():
    loss = torch.tensor(1.5)
    out = shim_mllm.FusedCausalLMOutput(loss=loss, predicted_ids=torch.zeros(1, 2, dtype=torch.int64), logits=None)
    assert out["loss"] is loss and out[0] is loss
    assert shim_mllm.FusedCausalLMOutput(loss=None, predicted_ids=torch.zeros(1, 2), logits="L")[0] == "L"


@pytest.mark.skipif(not NO_GPU, reason="checks the refusal on a machine without a GPU")
def test_shims_refuse_to_compute_without_a_gpu():
    """No CPU arithmetic path behind the reference's signatures: with CPU tensors and no CUDA device
    every compute shim raises (the oracle is test infrastructure, never a fallback)."""
    tok = ToyTokenizer(100)
    table = torch.randn(100, 16).to(torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        shim_img.average_embeddings_for_tokens(tok, {"initial": table}, ["red", "sky blue"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        shim_mllm.lm_head_loss_and_argmax(torch.randn(1, 4, 16).to(torch.bfloat16), table,
                                          torch.tensor([[-100, 3, 5, -100]]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        shim_tea.pairwise_cosine_similarity(table[:6].float())
    from multimodal_concept_learning_b200.shims.vision_training import classifier_loss_and_top1
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        classifier_loss_and_top1(torch.randn(4, 16), torch.randn(10, 16), torch.zeros(10), torch.tensor([1, 2, 3, 4]))


def test_lazy_logits_surface_and_dispatch_without_touching_the_operands():
    """``outputs.logits`` under the fused forwards: tensor-like shape / dtype / ``.data`` without any
    computation; the three reads the reference makes (argmax, max(.,1), cross_entropy) are routed to
    the fused scan -- on a machine without a GPU they must raise, not fall back to a dense matmul."""
    from multimodal_concept_learning_b200.shims.lazy_logits import LazyLogits
    f = torch.randn(2, 5, 16).to(torch.bfloat16)
    w = torch.randn(40, 16).to(torch.bfloat16)
    lazy = LazyLogits(f, w)
    assert lazy.shape == torch.Size((2, 5, 40)) and lazy.size(-1) == 40 and lazy.dim() == 3 and len(lazy) == 2
    assert lazy.dtype == torch.bfloat16 and lazy.device == f.device
    assert lazy.data is lazy and lazy.detach() is lazy and lazy._dense is None
    assert "materialized=False" in repr(lazy)
    if NO_GPU:
        for read in (lambda: torch.argmax(lazy, dim=-1), lambda: torch.max(lazy.data, 2),
                     lambda: LazyLogits(f[0], w).cross_entropy(torch.tensor([1, 2, 3, 4, 5]))):
            with pytest.raises(RuntimeError, match="no CPU fallback"):
                read()
        assert lazy._dense is None, "a refused fused read must not have materialised the logits"
    # any OTHER use is the reference's own expression (nn.Linear), computed once and cached
    dense = lazy.float()
    torch.testing.assert_close(dense, (f.float() @ w.float().T).to(torch.bfloat16).float(), rtol=2e-2, atol=2e-2)
    assert lazy._dense is not None and lazy[0].shape == (5, 40)
