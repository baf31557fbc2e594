"""GPU: the drop-in shims (reference function names) against the golden vectors produced by
the reference's own functions, and against the oracle restatements."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import concept_scan_ref as R
from oracle import reference_sites as S
from oracle.gen_golden import ToyTokenizer

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _gold(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


@pytest.fixture(scope="module", autouse=True)
def _built(lib_built):
    return lib_built


def test_a1_color_correlation_shim_matches_reference_golden(capsys):
    from multimodal_concept_learning_b200.shims.token_embedding_analysis import (
        calculate_color_embedding_correlation, pairwise_cosine_similarity)
    g = _gold("a1_color_correlation.npz")
    mapping = json.load(open(os.path.join(GOLD, str(g["mapping_name"]))))
    ood = [v for v in mapping.values() if v.startswith("<ood")]
    reg = [v for v in mapping.values() if not v.startswith("<ood")]
    t0 = torch.from_numpy(g["table_initial"]).to(torch.bfloat16)
    t3 = torch.from_numpy(g["table_epoch3"]).to(torch.bfloat16)
    emb = {"initial": t0, "epoch_0": t0.clone(), "epoch_3": t3}
    ood_ids, reg_ids = g["ood_ids"].tolist(), g["reg_ids"].tolist()
    r = calculate_color_embedding_correlation(emb, ood, reg, ood_ids, reg_ids, mapping)
    # 1 - cos is ill-conditioned on this anisotropic table (distances 3e-4..7e-4): an fp32
    # rounding of cos (1e-7) moves r in the 5th digit, in the reference's own fp32 run as well
    assert isinstance(r, float) and abs(r - float(g["r_last"])) < 5e-5
    assert "Pearson correlation coefficient" in capsys.readouterr().out      # same prints
    r0 = calculate_color_embedding_correlation({"initial": t0}, ood, reg, ood_ids, reg_ids, mapping)
    assert abs(r0 - float(g["r_initial_only"])) < 5e-5
    # the matrix itself against sklearn, incl. a zero row and the duplicated OOD rows
    from sklearn.metrics.pairwise import cosine_similarity
    e = t0[ood_ids + reg_ids].clone()
    e[1] = 0
    cos = pairwise_cosine_similarity(e).cpu().numpy()
    np.testing.assert_allclose(cos, cosine_similarity(e.float().numpy()), atol=2e-6)


def test_a1_nearest_tokens_config1(capsys):
    from multimodal_concept_learning_b200.shims.token_embedding_analysis import nearest_tokens
    from tests.util import check_topk
    g = torch.Generator().manual_seed(40)
    table = torch.randn(50257, 768, generator=g).to(torch.bfloat16)
    ids = torch.randperm(50257, generator=g)[:16].tolist()
    val, idx = nearest_tokens(table.cuda(), ids, k=50)
    ref = R.concept_scan_ref(table[ids], table, 50, keep_scores=True)
    check_topk(val, idx, ref.scores, 50, rtol=1e-4, atol=1e-5)
    assert idx[:, 0].cpu().tolist() == ids                    # each token is its own nearest neighbour


def test_a3_average_embeddings_shim_matches_reference_golden():
    from multimodal_concept_learning_b200.shims.token_embedding_analysis_imagenet import \
        average_embeddings_for_tokens
    from multimodal_concept_learning_b200.shims.multi_token import (get_averaged_embedding,
                                                                     get_averaged_embeddings)
    g = _gold("a3_average_embeddings.npz")
    tok = ToyTokenizer(int(g["V"]))
    names = [str(n) for n in g["names"]]
    t_bf = torch.from_numpy(g["table_initial"]).to(torch.bfloat16)
    t_f32 = torch.from_numpy(g["table_epoch0"])
    res = average_embeddings_for_tokens(tok, {"initial": t_bf, "epoch_0": t_f32}, names)
    assert res["initial"].dtype == torch.bfloat16 and not res["initial"].is_cuda     # same type/device out
    assert torch.equal(res["initial"].float(), torch.from_numpy(g["out_initial"]))   # bit-exact
    np.testing.assert_allclose(res["epoch_0"].numpy(), g["out_epoch0"], rtol=3e-7, atol=1e-7)
    assert average_embeddings_for_tokens(tok, {}, names) == {}
    assert average_embeddings_for_tokens(tok, {"initial": t_bf}, [])["initial"].shape == (0, t_bf.shape[1])
    # notebook twin, CUDA-resident table
    one = get_averaged_embedding(names[1], tok, t_bf.cuda())
    assert torch.equal(one.cpu().float(), torch.from_numpy(g["out_initial"])[1])
    many = get_averaged_embeddings([n for n in names if n], tok, t_f32.cuda(), normalize=True)
    want = torch.from_numpy(g["out_epoch0"])[[i for i, n in enumerate(names) if n]]
    want = want / want.norm(dim=1, keepdim=True)              # notebook cell 3 line 16
    torch.testing.assert_close(many.cpu(), want, rtol=1e-5, atol=1e-6)


def test_a4_a5_lm_head_shim_matches_reference_golden():
    from multimodal_concept_learning_b200.shims.mllm import lm_head_loss_and_argmax
    from multimodal_concept_learning_b200.shims.multimodal_training import (count_yes_no_matches,
                                                                             evaluate_hidden_batches)
    g = _gold("a4_a5_mllm_head.npz")
    hidden = torch.from_numpy(g["hidden"]).to(torch.bfloat16)
    table = torch.from_numpy(g["table"]).to(torch.bfloat16)
    labels = torch.from_numpy(g["labels"])
    ref_logits = torch.from_numpy(g["logits"])                # the reference's bf16 logits
    tok = ToyTokenizer(table.shape[0], int(g["yes_id"]), int(g["no_id"]))
    for rows in ("labelled", "all"):
        out = lm_head_loss_and_argmax(hidden.cuda(), table.cuda(), labels.cuda(), rows=rows)
        # bf16 bound of north_star (the reference rounds its logits to bf16 before the CE)
        assert abs(float(out.loss) - float(g["loss"])) <= 1e-2 * abs(float(g["loss"]))
        # tighter: against the exact scores the kernel is meant to compute
        exact, _ = S.causal_lm_head_loss_ref(hidden.double(), table.double(), labels, logits_dtype=torch.float64)
        assert abs(float(out.loss) - float(exact)) <= 1e-4 * abs(float(exact))
        pred = out.predicted_ids.cpu()
        mask = labels != -100 if rows == "labelled" else torch.ones_like(labels, dtype=torch.bool)
        ref_pred = ref_logits.argmax(-1)
        differs = (pred != ref_pred) & mask
        # where the fused argmax differs from the bf16-logit argmax it must be a bf16 near-tie
        z = (hidden.double().reshape(-1, hidden.shape[-1]) @ table.double().T).reshape(ref_logits.shape)
        for b, t in differs.nonzero().tolist():
            assert z[b, t, pred[b, t]] >= z[b, t, ref_pred[b, t]] - 1e-12
        c, n = count_yes_no_matches(pred, labels, tok)
        assert n == 2
    ev = evaluate_hidden_batches([{"hidden_states": hidden.cuda(), "labels": labels.cuda()}], table.cuda(), tok)
    assert set(ev) == {"test_loss", "test_acc"}
    assert abs(ev["test_loss"] - float(g["test_loss"])) <= 1e-2 * abs(float(g["test_loss"]))
    exact_pred = z.argmax(-1)
    c_exact, n_exact = S.evaluate_predictions_ref(z, labels, tok)[:2]
    assert abs(ev["test_acc"] - 100.0 * c_exact / n_exact) < 1e-9


def test_a7_vision_head_shim_matches_golden():
    from multimodal_concept_learning_b200.shims.vision_training import classifier_loss_and_top1
    g = _gold("a7_vision_head.npz")
    feats, w = torch.from_numpy(g["feats"]), torch.from_numpy(g["weight"])
    labels = torch.from_numpy(g["labels"])
    for eps in (0.0, 0.1):
        loss, pred = classifier_loss_and_top1(feats.cuda(), w.cuda(), None, labels.cuda(), eps)
        assert abs(float(loss) - float(g[f"loss_{eps}"])) <= 1e-4 * float(g[f"loss_{eps}"])
        assert torch.equal(pred.cpu(), torch.from_numpy(g["predicted"]))         # first max wins
    # with a bias and fp16 features (the vision launch script uses fp16 autocast)
    bias = torch.linspace(-1, 1, w.shape[0])
    loss, pred = classifier_loss_and_top1(feats.half().cuda(), w.half().cuda(), bias.cuda(), labels.cuda(), 0.1)
    want, wpred, _ = S.vision_ce_top1_ref(feats.half().float(), w.half().float(), bias, labels, 0.1)
    assert abs(float(loss) - float(want)) <= 1e-4 * float(want)
    assert torch.equal(pred.cpu(), wpred)


@pytest.mark.parametrize("dtype,rtol", [(torch.float32, 1e-4), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("eps", [0.0, 0.1])
def test_f1_fused_cross_entropy_backward_matches_autograd(dtype, rtol, eps):
    """gradients of the fused CE w.r.t. hidden states and the (tied) table against
    torch autograd through the reference formulation F.cross_entropy(h @ E^T)."""
    from multimodal_concept_learning_b200.autograd import fused_cross_entropy
    g = torch.Generator().manual_seed(50)
    Q, V, D = 96, 3001, 64
    h = (torch.randn(Q, D, generator=g) * 0.5).to(dtype).cuda().requires_grad_(True)
    E = (torch.randn(V, D, generator=g) * 0.5).to(dtype).cuda().requires_grad_(True)
    labels = torch.randint(0, V, (Q,), generator=g)
    labels[::3] = -100
    labels = labels.cuda()
    loss, pred = fused_cross_entropy(h, E, labels, label_smoothing=eps, chunk_rows=1000)
    loss.backward()
    h2 = h.detach().double().requires_grad_(True)
    E2 = E.detach().double().requires_grad_(True)
    ref = F.cross_entropy(h2 @ E2.T, labels, ignore_index=-100, label_smoothing=eps)
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 1e-4 * abs(float(ref))
    scale_h, scale_E = h2.grad.abs().max(), E2.grad.abs().max()
    assert (h.grad.double() - h2.grad).abs().max() <= rtol * scale_h
    assert (E.grad.double() - E2.grad).abs().max() <= rtol * scale_E
    assert (h.grad[::3] == 0).all()                       # ignored rows get no gradient
    assert torch.equal(pred.cpu(), (h2 @ E2.T).argmax(-1).cpu()) or dtype == torch.bfloat16


def test_f1_lm_head_shim_trains_the_table():
    """`language_embed_only` (mllm.py:181-184): the tied table receives gradients through the
    shim exactly as through HF's logits + ForCausalLMLoss."""
    from multimodal_concept_learning_b200.shims.mllm import lm_head_loss_and_argmax
    g = torch.Generator().manual_seed(60)
    B, T, D, V = 3, 9, 32, 700
    hidden = (torch.randn(B, T, D, generator=g) * 0.3).cuda()
    table = (torch.randn(V, D, generator=g) * 0.3).cuda().requires_grad_(True)
    labels = torch.full((B, T), -100, dtype=torch.long)
    labels[0, 5:7] = torch.tensor([3, 699])
    labels[2, 8] = 41
    out = lm_head_loss_and_argmax(hidden, table, labels.cuda())
    out.loss.backward()
    t2 = table.detach().double().requires_grad_(True)
    want, _ = S.causal_lm_head_loss_ref(hidden.double().cpu(), t2.cpu(), labels, logits_dtype=torch.float64)
    want.backward()
    assert abs(float(out.loss.detach()) - float(want.detach())) <= 1e-4 * abs(float(want.detach()))
    assert (table.grad.cpu().double() - t2.grad.cpu()).abs().max() <= 1e-4 * t2.grad.abs().max()
    with torch.no_grad():
        ev = lm_head_loss_and_argmax(hidden, table, labels.cuda())
    assert not ev.loss.requires_grad


def test_host_query_pipeline_reused_results_live_lag_plus_one_yields():
    """with reuse_host_buffers=True the results rotate through 2*lag + 2 pinned sets; a
    yielded result must stay intact until lag + 1 MORE results have been yielded (no copy may land
    in it earlier).  Every result is held un-cloned next to a clone and compared after lag + 1
    further yields, with batches that all give different answers."""
    from multimodal_concept_learning_b200.pipeline import HostQueryPipeline
    g = torch.Generator().manual_seed(71)
    table = torch.randn(4000, 64, generator=g).to(torch.bfloat16).cuda()
    batches = [torch.randn(150, 64, generator=g).to(torch.bfloat16).pin_memory() for _ in range(14)]
    lag = 2
    pipe = HostQueryPipeline(table, 10, scale=10.0, lag=lag, reuse_host_buffers=True)
    held = []                                             # (live tensors, clones) in yield order
    checked = 0
    for res in pipe.run(batches):
        held.append((res, tuple(t.clone() for t in res)))
        if len(held) > lag + 1:
            live, snap = held[len(held) - 1 - (lag + 1)]  # lag + 1 results have been yielded since
            torch.cuda.synchronize()                      # any copy that could touch it has landed
            for a, b in zip(live, snap):
                assert torch.equal(a, b), "a reused host buffer was overwritten inside its documented lifetime"
            checked += 1
    assert checked == len(batches) - (lag + 1)
    assert not torch.equal(held[0][1][1], held[1][1][1])  # the batches really differ


@pytest.mark.parametrize("reuse", [False, True])
def test_host_query_pipeline_matches_direct_scan(reuse):
    """Fresh pinned results per batch, or a ring of lag + 2 reused result sets (each result is
    then checked as it is yielded, before the ring wraps)."""
    import multimodal_concept_learning_b200 as mcl
    from multimodal_concept_learning_b200.pipeline import HostQueryPipeline
    g = torch.Generator().manual_seed(70)
    table = torch.randn(5000, 128, generator=g).to(torch.bfloat16).cuda()
    batches = [torch.randn(200, 128, generator=g).to(torch.bfloat16).pin_memory() for _ in range(9)]
    refs = [mcl.concept_scan(b.cuda(), table, 20, scale=10.0) for b in batches]
    pipe = HostQueryPipeline(table, 20, scale=10.0, lag=2, reuse_host_buffers=reuse)
    n = 0
    for ref, (val, idx, stats) in zip(refs, pipe.run(batches)):
        assert not val.is_cuda and torch.equal(val, ref.topk_val.cpu()) and torch.equal(idx, ref.topk_idx.cpu())
        torch.testing.assert_close(stats, ref.stats.cpu(), rtol=1e-6, atol=1e-6)
        n += 1
    assert n == 9


# ---- round 2: the documented patches, end to end on tiny models -----------------------------
def _load_state(module, npz, prefix="sd::"):
    sd = {}
    for key in npz.files:
        if key.startswith(prefix) and not key.endswith("::bf16"):
            t = torch.from_numpy(npz[key])
            if bool(npz[key + "::bf16"]):
                t = t.to(torch.bfloat16)
            sd[key[len(prefix):]] = t
    missing, unexpected = module.load_state_dict(sd, strict=False)
    assert not unexpected and all("lm_head" in k for k in missing), (missing, unexpected)   # tied head


class _Accelerator:
    """The Accelerator surface `evaluate_model` touches (multimodal_training.py:258,264,309)."""
    is_main_process = True

    def unwrap_model(self, model):
        return model

    def autocast(self):
        return torch.autocast("cuda", dtype=torch.bfloat16)


def test_a4_a5_patched_mllm_forward_and_evaluate_model_reproduce_the_reference(capsys):
    """INTEGRATION.md's patch, executed: an MLLM-shaped model (the reference's attribute names and
    state-dict keys, golden weights of the tiny Gemma-3 + ViT the reference's own forward was run
    on) gets `forward = shims.mllm.fused_forward`; `outputs.loss`, `torch.argmax(outputs.logits)`
    and the drop-in `evaluate_model(model, test_loader, config, accelerator)` must reproduce the
    reference's loss / test_loss / test_acc.  The decoder runs in bf16 on the GPU here and on the
    CPU in the golden run, so the bound is north_star's bf16 one (rtol 1e-2)."""
    import types
    from tests.tiny_models import TinyMLLM
    from multimodal_concept_learning_b200.shims.lazy_logits import LazyLogits
    from multimodal_concept_learning_b200.shims.mllm import fused_forward
    from multimodal_concept_learning_b200.shims.multimodal_training import evaluate_model
    gm, gh = _gold("a4_a5_mllm_model.npz"), _gold("a4_a5_mllm_head.npz")
    model = TinyMLLM()
    _load_state(model, gm)
    model.forward = types.MethodType(fused_forward, model)       # == MLLM.forward = fused_forward
    model = model.cuda().eval()
    batch = {"images": torch.from_numpy(gm["images"]).cuda(), "input_ids": torch.from_numpy(gm["input_ids"]).cuda(),
             "attention_mask": torch.from_numpy(gm["attention_mask"]).cuda(),
             "labels": torch.from_numpy(gh["labels"]).cuda()}
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        out = model(**batch)
    assert abs(float(out.loss) - float(gh["loss"])) <= 1e-2 * abs(float(gh["loss"]))
    assert isinstance(out.logits, LazyLogits) and tuple(out.logits.shape) == tuple(gh["logits"].shape)
    assert out.logits._dense is None, "the loss must not have materialised the logits"
    # multimodal_training.py:276 verbatim; answered by the fused k=1 scan, not by a [B,T,V] tensor
    predicted_ids = torch.argmax(out.logits, dim=-1)
    assert out.logits._dense is None and predicted_ids.shape == batch["labels"].shape
    ref_logits = torch.from_numpy(gh["logits"])
    ref_pred = ref_logits.argmax(-1)
    differs = (predicted_ids.cpu() != ref_pred)
    # the decoders differ in the last bf16 bits (GPU vs CPU kernels): a different argmax must be a near-tie
    top2 = ref_logits.float().topk(2, dim=-1).values
    assert (~differs | ((top2[..., 0] - top2[..., 1]) < 0.15)).all()
    # any other use of .logits materialises the real tensor, close to the reference's bf16 logits
    dense = out.logits.float()
    assert out.logits._dense is not None and dense.shape == ref_logits.shape
    torch.testing.assert_close(dense.cpu(), ref_logits.float(), rtol=5e-2, atol=0.3)
    # a5: reference signature, reference prints, reference metrics
    cfg = types.SimpleNamespace(disable_tqdm=True)
    ev = evaluate_model(model, [batch], cfg, _Accelerator())
    assert set(ev) == {"test_loss", "test_acc"}
    assert abs(ev["test_loss"] - float(gh["test_loss"])) <= 1e-2 * abs(float(gh["test_loss"]))
    assert ev["test_acc"] == float(gh["test_acc"])
    assert "Test Accuracy" in capsys.readouterr().out
    # training mode: the same forward is differentiable down to the table (language_embed_only)
    model.train()
    for p in model.parameters():
        p.requires_grad_(False)
    table = model.language_model.get_input_embeddings().weight
    table.requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = model(**batch).loss
    loss.backward()
    assert table.grad is not None and torch.isfinite(table.grad).all() and float(table.grad.abs().sum()) > 0


@pytest.mark.parametrize("eps", [0.0, 0.1])
def test_a7_patched_vit_forward_and_criterion_reproduce_the_reference(eps):
    """vision_training.py:81-83,115-116,132 with the documented patch: ViTForImageClassification.forward
    = fused_vit_forward, criterion = FusedCrossEntropyAndTop1(label_smoothing); golden weights, images,
    loss and predictions come from the stock HF forward + nn.CrossEntropyLoss + torch.max."""
    import types
    from transformers import ViTForImageClassification
    from tests.tiny_models import tiny_vit_config
    from multimodal_concept_learning_b200.shims.vision_training import FusedCrossEntropyAndTop1, fused_vit_forward
    g = _gold("a7_vit_model.npz")
    model = ViTForImageClassification(tiny_vit_config(num_labels=int(g["num_labels"])))
    _load_state(model, g)
    model.forward = types.MethodType(fused_vit_forward, model)
    model = model.cuda().eval()
    images, labels = torch.from_numpy(g["images"]).cuda(), torch.from_numpy(g["labels"]).cuda()
    criterion = FusedCrossEntropyAndTop1(label_smoothing=eps)
    with torch.no_grad():
        outputs = model(images)
        loss = criterion(outputs.logits, labels)                 # :116
        _, predicted = torch.max(outputs.logits.data, 1)         # :132
    assert outputs.logits._dense is None
    assert abs(float(loss) - float(g[f"loss_{eps}"])) <= 1e-4 * abs(float(g[f"loss_{eps}"]))
    assert torch.equal(predicted.cpu(), torch.from_numpy(g["predicted"]))
    # stock nn.CrossEntropyLoss dispatches to the fused path too (F.cross_entropy on a LazyLogits)
    stock = torch.nn.CrossEntropyLoss(label_smoothing=eps)(outputs.logits, labels)
    assert outputs.logits._dense is None
    torch.testing.assert_close(stock, loss, rtol=1e-6, atol=1e-6)
    # training: gradients reach the classifier weight and bias and the encoder
    model.train()
    out = model(images)
    criterion(out.logits, labels).backward()
    ref = ViTForImageClassification(tiny_vit_config(num_labels=int(g["num_labels"])))
    _load_state(ref, g)
    ref = ref.cuda().train()
    torch.nn.CrossEntropyLoss(label_smoothing=eps)(ref(images).logits, labels).backward()
    torch.testing.assert_close(model.classifier.weight.grad, ref.classifier.weight.grad, rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(model.classifier.bias.grad, ref.classifier.bias.grad, rtol=1e-3, atol=1e-5)
    gw, rw = model.vit.embeddings.cls_token.grad, ref.vit.embeddings.cls_token.grad
    torch.testing.assert_close(gw, rw, rtol=1e-2, atol=1e-5)


# ---- round 2: native tcgen05 backward -------------------------------------------------------
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 520, 200), (1000, 72, 1032), (77, 1152, 4104)])
def test_gemm_bf16_all_operand_layouts(a_mn, b_mn, M, N, K):
    """The backward's GEMM on its own: K-major and MN-major operand descriptors (the latter are what
    make dL/dq = P T and dL/dT = P^T q transposition-free), ragged M / N / K, accumulate on / off."""
    from multimodal_concept_learning_b200 import _lib, ops
    lib = _lib.load()
    g = torch.Generator().manual_seed(M + N + K)
    pad = lambda x: -(-x // 8) * 8                                   # pitches: multiples of 16 bytes
    A = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
    B = torch.randn(K, N, generator=g).to(torch.bfloat16).cuda()
    want = A.float() @ B.float()
    a_store = torch.zeros((K, pad(M)) if a_mn else (M, pad(K)), dtype=torch.bfloat16, device="cuda")
    b_store = torch.zeros((K, pad(N)) if b_mn else (N, pad(K)), dtype=torch.bfloat16, device="cuda")
    (a_store[:, :M] if a_mn else a_store[:, :K]).copy_(A.t() if a_mn else A)
    (b_store[:, :N] if b_mn else b_store[:, :K]).copy_(B if b_mn else B.t())
    C = torch.full((M, N), 3.0, dtype=torch.float32, device="cuda")
    for accumulate, expect in ((0, want), (1, 2 * want)):
        _lib.check(lib.mcl_gemm_bf16(a_store.data_ptr(), a_mn, a_store.stride(0), b_store.data_ptr(), b_mn,
                                     b_store.stride(0), C.data_ptr(), C.stride(0), M, N, K, accumulate,
                                     ops._stream(C.device)))
        torch.cuda.synchronize()
        torch.testing.assert_close(C, expect, rtol=1e-4, atol=1e-3 * K ** 0.5)


@pytest.mark.parametrize("Q,V,D,eps,cap", [(700, 9000, 128, 0.1, None), (300, 50000, 64, 0.0, None),
                                           (5000, 20000, 64, 0.1, None), (260, 3000, 72, 0.0, 6.0),
                                           (24, 262235, 1152, 0.0, None), (9000, 6000, 64, 0.0, None)])
def test_f1_native_backward_blocks_and_chunks(Q, V, D, eps, cap):
    """bf16 backward on the tensor cores across the library's blocking: several row blocks, several
    table chunks (V beyond one dL/dz block), several blocks of 4096 rows, soft-capped logits, and the
    reference's own shape (a few labelled rows against the Gemma-3 table: dL/dq with its K range split
    over the SMs, the table gradient written in bf16 by the GEMM), 6750 labelled rows (two row blocks:
    the fp32 table gradient accumulates across them).  Reference: torch autograd
    in fp32 through F.cross_entropy(softcap(h @ E^T)) on the same bf16 values."""
    from multimodal_concept_learning_b200.autograd import fused_cross_entropy
    g = torch.Generator(device="cuda").manual_seed(Q + V)
    h = (torch.randn(Q, D, generator=g, device="cuda") * 0.4).to(torch.bfloat16).requires_grad_(True)
    E = (torch.randn(V, D, generator=g, device="cuda") * 0.4).to(torch.bfloat16).requires_grad_(True)
    labels = torch.randint(0, V, (Q,), generator=g, device="cuda")
    labels[::4] = -100
    loss, pred = fused_cross_entropy(h, E, labels, label_smoothing=eps, softcap=cap)
    (loss * 3.0).backward()                                  # a non-trivial upstream gradient
    h2 = h.detach().float().requires_grad_(True)
    E2 = E.detach().float().requires_grad_(True)
    z = h2 @ E2.T
    if cap:
        z = torch.tanh(z / cap) * cap
    ref = F.cross_entropy(z, labels, ignore_index=-100, label_smoothing=eps)
    (ref * 3.0).backward()
    assert abs(float(loss) - float(ref)) <= 1e-4 * abs(float(ref))
    for got, want in ((h.grad, h2.grad), (E.grad, E2.grad)):
        scale = want.abs().max()
        assert (got.float() - want).abs().max() <= 1e-2 * scale, float((got.float() - want).abs().max() / scale)
    assert (h.grad[::4] == 0).all()


def test_f1_backward_uses_no_library_gemm():
    """'autograd.py contains no torch.mm' -- the backward is the library's own kernels."""
    import inspect
    import multimodal_concept_learning_b200.autograd as ag
    src = inspect.getsource(ag)
    for banned in ("torch.mm", "torch.matmul", " @ ", "F.linear", "einsum", "torch.bmm"):
        assert banned not in src, banned
    import multimodal_concept_learning_b200 as mcl
    from multimodal_concept_learning_b200.autograd import fused_cross_entropy
    h = torch.randn(200, 64, device="cuda").bfloat16().requires_grad_(True)
    E = torch.randn(3000, 64, device="cuda").bfloat16().requires_grad_(True)
    loss, _ = fused_cross_entropy(h, E, torch.randint(0, 3000, (200,), device="cuda"))
    n0 = mcl.launch_count()
    loss.backward()
    assert mcl.launch_count() - n0 == 3, "dL/dz (scan kernel, grad epilogue) + two tcgen05 GEMMs"
