"""CPU: pin the oracle against (a) outputs of the REFERENCE's own functions, committed
under tests/golden/ by oracle/gen_golden.py, and (b) the third-party primitives the
reference calls (sklearn, HF ForCausalLMLoss, nn.CrossEntropyLoss, torch.topk/argmax)."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import concept_scan_ref as R
from oracle import reference_sites as S
from oracle.gen_golden import ToyTokenizer

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _gold(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def _color_mapping(name):
    # the mapping json is reference data; a copy of the one used is kept beside the vectors
    with open(os.path.join(GOLD, str(name))) as f:
        return json.load(f)


# ---- a1 -------------------------------------------------------------------------------
def test_a1_color_correlation_matches_reference_run():
    g = _gold("a1_color_correlation.npz")
    mapping = _color_mapping(g["mapping_name"])
    ood = [v for v in mapping.values() if v.startswith("<ood")]
    reg = [v for v in mapping.values() if not v.startswith("<ood")]
    t0 = torch.from_numpy(g["table_initial"]).to(torch.bfloat16)
    t3 = torch.from_numpy(g["table_epoch3"]).to(torch.bfloat16)
    emb = {"initial": t0, "epoch_0": t0.clone(), "epoch_3": t3}
    ood_ids, reg_ids = g["ood_ids"].tolist(), g["reg_ids"].tolist()
    for tag, e in (("last", emb), ("initial_only", {"initial": t0})):
        want = float(g[f"r_{tag}"])
        r_loop, cd, ed = S.color_embedding_correlation_loop_ref(e, ood, reg, ood_ids, reg_ids, mapping)
        r_bat, cd2, ed2 = S.color_embedding_correlation_batched_ref(e, ood, reg, ood_ids, reg_ids, mapping)
        assert abs(r_loop - want) < 1e-6, (tag, r_loop, want)      # literal restatement
        assert abs(r_bat - want) < 1e-5, (tag, r_bat, want)        # batched formulation
        np.testing.assert_allclose(cd, cd2, atol=1e-12)
        np.testing.assert_allclose(ed, ed2, atol=2e-6)             # fp32 pairwise vs fp64 matrix


def test_a1_pair_restatement_equals_sklearn():
    from sklearn.metrics.pairwise import cosine_similarity
    g = torch.Generator().manual_seed(5)
    e = torch.randn(9, 33, generator=g).numpy().astype(np.float32)
    e[4] = 0
    for i in range(9):
        for j in range(9):
            want = cosine_similarity([e[i]], [e[j]])[0][0]
            got = S._cosine_pair_sklearn_semantics(e[i], e[j])
            assert abs(got - want) <= 2e-7


def test_sklearn_docstring_vector_zero_rows():
    # sklearn/metrics/pairwise.py docstring: zero rows give similarity 0
    x = torch.tensor([[0., 0., 0.], [1., 1., 1.]])
    y = torch.tensor([[1., 0., 0.], [1., 1., 0.]])
    z = R.scores_ref(x, y)
    np.testing.assert_allclose(z.numpy(), [[0, 0], [0.57735027, 0.81649658]], atol=1e-7)


# ---- a3 -------------------------------------------------------------------------------
def test_a3_average_embeddings_matches_reference_run():
    g = _gold("a3_average_embeddings.npz")
    V = int(g["V"])
    tok = ToyTokenizer(V)
    names = [str(n) for n in g["names"]]
    t_bf = torch.from_numpy(g["table_initial"]).to(torch.bfloat16)
    t_f32 = torch.from_numpy(g["table_epoch0"])
    res = S.average_embeddings_for_tokens_ref(tok, {"initial": t_bf, "epoch_0": t_f32}, names)
    assert bool(g["out_initial_is_bf16"]) and res["initial"].dtype == torch.bfloat16
    assert torch.equal(res["initial"].float(), torch.from_numpy(g["out_initial"]))
    np.testing.assert_allclose(res["epoch_0"].numpy(), g["out_epoch0"], rtol=0, atol=0)
    # CSR formulation used by the CUDA path: bit-exact for bf16 tables, 1 ulp for fp32
    ids, offs = [], [0]
    for n in names:
        ids += tok.encode(n)
        offs.append(len(ids))
    csr_bf = R.gather_mean_ref(t_bf, offs, ids)
    assert torch.equal(csr_bf.float(), torch.from_numpy(g["out_initial"]))
    csr_f32 = R.gather_mean_ref(t_f32, offs, ids)
    np.testing.assert_allclose(csr_f32.numpy(), g["out_epoch0"], rtol=3e-7, atol=1e-7)
    assert (csr_bf[3] == 0).all()      # the empty name -> zeros (imagenet.py:283)


# ---- a4 / a5 --------------------------------------------------------------------------
def test_a4_lm_head_loss_matches_reference_run():
    g = _gold("a4_a5_mllm_head.npz")
    assert float(g["softcap"]) == 0.0      # Gemma-3: final_logit_softcapping is None
    hidden = torch.from_numpy(g["hidden"]).to(torch.bfloat16)
    table = torch.from_numpy(g["table"]).to(torch.bfloat16)
    labels = torch.from_numpy(g["labels"])
    loss, logits = S.causal_lm_head_loss_ref(hidden, table, labels, logits_dtype=torch.bfloat16)
    # same bf16 GEMM, same loss
    assert torch.equal(logits.float(), torch.from_numpy(g["logits"]))
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    # the fused formulation: only rows with a shifted label, fp64 ground truth
    B, T, D = hidden.shape
    shift = F.pad(labels, (0, 1), value=-100)[..., 1:].reshape(-1)
    r = R.concept_scan_ref(hidden.reshape(B * T, D), table, 1, normalize_q=False,
                           normalize_t=False, labels=shift)
    # bf16 logits rounding in the reference vs exact scores here: rtol 1e-2 (north_star)
    assert abs(float(r.loss) - float(g["loss"])) <= 1e-2 * abs(float(g["loss"]))


def test_a5_evaluate_model_matches_reference_run():
    g = _gold("a4_a5_mllm_head.npz")
    logits = torch.from_numpy(g["logits"])
    labels = torch.from_numpy(g["labels"])
    tok = ToyTokenizer(logits.shape[-1], int(g["yes_id"]), int(g["no_id"]))
    correct, total, pred = S.evaluate_predictions_ref(logits, labels, tok)
    assert total == 2                                   # sample 2 has no label -> skipped
    assert abs(100.0 * correct / total - float(g["test_acc"])) < 1e-9
    assert abs(float(g["test_loss"]) - float(g["loss"])) < 1e-9


# ---- a7 -------------------------------------------------------------------------------
def test_a7_vision_head_matches_nn_modules():
    g = _gold("a7_vision_head.npz")
    feats, w = torch.from_numpy(g["feats"]), torch.from_numpy(g["weight"])
    labels = torch.from_numpy(g["labels"])
    for eps in (0.0, 0.1):
        loss, pred, _ = S.vision_ce_top1_ref(feats, w, None, labels, eps)
        assert abs(float(loss) - float(g[f"loss_{eps}"])) < 1e-6
        assert torch.equal(pred, torch.from_numpy(g["predicted"]))
        r = R.concept_scan_ref(feats, w, 1, normalize_q=False, normalize_t=False, labels=labels,
                               label_smoothing=eps)
        assert abs(float(r.loss) - float(g[f"loss_{eps}"])) < 1e-5
        # first-max-wins == lowest index wins (row C-1 duplicates row 0)
        assert torch.equal(r.topk_idx[:, 0], torch.from_numpy(g["predicted"]))


# ---- the composed oracle vs library primitives ----------------------------------------
@pytest.mark.parametrize("normalize,scale,eps", [(True, 1.0, 0.0), (True, 100.0, 0.1), (False, 1.0, 0.0)])
def test_concept_scan_ref_vs_torch_composition(normalize, scale, eps):
    g = torch.Generator().manual_seed(11)
    q, t = torch.randn(37, 48, generator=g), torch.randn(301, 48, generator=g)
    labels = torch.randint(0, 301, (37,), generator=g)
    labels[::5] = -100
    r = R.concept_scan_ref(q, t, 7, normalize_q=normalize, normalize_t=normalize, scale=scale,
                           labels=labels, label_smoothing=eps)
    c = R.torch_composition_ref(q, t, 7, normalize=normalize, scale=scale, labels=labels,
                                label_smoothing=eps)
    torch.testing.assert_close(r.topk_val.float(), c["topk_val"], rtol=1e-4, atol=1e-5)
    assert torch.equal(r.topk_idx, c["topk_idx"])      # random data: no ties
    torch.testing.assert_close(r.lse.float(), c["lse"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(r.loss.float(), c["loss"], rtol=1e-5, atol=1e-5)


def test_topk_tie_rule_lowest_index():
    z = torch.tensor([[1., 3., 3., 2., 3.]])
    v, i = R.topk_lowest_index(z, 3)
    assert i.tolist() == [[1, 2, 4]] and v.tolist() == [[3., 3., 3.]]
    assert int(torch.argmax(z, dim=-1)) == 1        # reference's argmax: first occurrence


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_sharded_oracle_equals_unsharded(world):
    g = torch.Generator().manual_seed(13)
    q, t = torch.randn(19, 32, generator=g), torch.randn(203, 32, generator=g)
    t[100:110] = t[:10]                                # cross-shard exact ties
    labels = torch.randint(0, 203, (19,), generator=g)
    labels[3] = -100
    a = R.concept_scan_ref(q, t, 12, labels=labels, label_smoothing=0.1, scale=10.0)
    b = R.concept_scan_sharded_ref(q, t, 12, world, labels=labels, label_smoothing=0.1, scale=10.0)
    assert torch.equal(a.topk_idx, b.topk_idx)
    torch.testing.assert_close(a.topk_val, b.topk_val, rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(a.lse, b.lse, rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(a.loss, b.loss, rtol=1e-12, atol=1e-12)


def test_experiment_label_maps_are_the_expected_fixtures():
    # the only real fixtures are the label maps (query-set sizes)
    m = _color_mapping(_gold("a1_color_correlation.npz")["mapping_name"])
    assert len(m) == 12 and sum(v.startswith("<ood") for v in m.values()) >= 1
    assert S.extract_rgb_from_mapping_ref(m, next(iter(m.values()))) != (0.5, 0.5, 0.5)
    assert S.extract_rgb_from_mapping_ref(m, "not-a-token") == (0.5, 0.5, 0.5)
