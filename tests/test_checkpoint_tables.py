"""CPU: the checkpoint-table reader finds the same tensors the reference's loader would
(``get_input_embeddings().weight`` per epoch, reference order, best_model.pt skipped) straight
from ``torch.save(state_dict)`` files."""
import os

import pytest
import torch

from multimodal_concept_learning_b200.shims.checkpoint_tables import (epoch_checkpoints, find_table_key,
                                                                       load_embeddings_by_epoch, load_table)


def _state_dict(seed, V=37, D=16):
    g = torch.Generator().manual_seed(seed)
    table = torch.randn(V, D, generator=g).to(torch.bfloat16)
    return {"vision_model.embeddings.cls_token": torch.zeros(1, 1, 8),
            "projector.weight": torch.randn(D, 8, generator=g),
            "language_model.model.embed_tokens.weight": table,
            "language_model.lm_head.weight": table,           # tied (modeling_gemma3.py:593)
            "language_model.model.layers.0.mlp.up_proj.weight": torch.randn(4, D, generator=g)}, table


def test_reads_every_epoch_in_reference_order(tmp_path):
    models = tmp_path / "models"
    models.mkdir()
    want = {}
    for name, seed in (("initial_model.pt", 0), ("epoch_10_model.pt", 10), ("epoch_2_model.pt", 2),
                       ("best_model.pt", 99)):
        sd, table = _state_dict(seed)
        torch.save(sd, models / name)
        want[name] = table
    files = epoch_checkpoints(str(models))
    assert list(files) == ["initial", "epoch_2", "epoch_10"]             # numeric sort, no best_model
    res = load_embeddings_by_epoch(str(tmp_path), device=None, verbose=False)
    assert list(res.tables) == ["initial", "epoch_2", "epoch_10"]
    assert torch.equal(res.tables["initial"], want["initial_model.pt"])
    assert torch.equal(res.tables["epoch_2"], want["epoch_2_model.pt"])
    assert torch.equal(res.tables["epoch_10"], want["epoch_10_model.pt"])
    assert res.tables["epoch_2"].dtype == torch.bfloat16 and not res.inv_norms
    assert torch.equal(load_table(str(models / "epoch_2_model.pt")), want["epoch_2_model.pt"])


def test_missing_pieces_raise(tmp_path):
    with pytest.raises(FileNotFoundError):
        load_embeddings_by_epoch(str(tmp_path), verbose=False)
    with pytest.raises(KeyError):
        find_table_key({"projector.weight": torch.zeros(1)})


@pytest.mark.gpu
def test_gpu_tables_cache_inverse_norms(tmp_path, lib_built):
    from oracle import concept_scan_ref as R
    from tests.util import check_topk
    models = tmp_path / "models"
    models.mkdir()
    sd, table = _state_dict(5, V=3000, D=64)
    torch.save(sd, models / "epoch_0_model.pt")
    res = load_embeddings_by_epoch(str(tmp_path), device="cuda", verbose=False)
    torch.testing.assert_close(res.inv_norms["epoch_0"].cpu().double(),
                               R.row_inv_norm_ref(table, torch.float64), rtol=1e-6, atol=0)
    q = table[:20].cuda()
    out = res.scan("epoch_0", q, 10)
    ref = R.concept_scan_ref(table[:20], table, 10, keep_scores=True)
    check_topk(out.topk_val, out.topk_idx, ref.scores, 10, rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
def test_gpu_all_epochs_in_one_graph_launch(tmp_path, lib_built):
    """SURVEY 8f-3, second half: the queries against every epoch's table with ONE launch from the
    host (a CUDA graph over the per-epoch scans); results equal the per-epoch scans, also when the
    graph is replayed with new queries."""
    models = tmp_path / "models"
    models.mkdir()
    tables = {}
    for name, seed in (("initial_model.pt", 1), ("epoch_0_model.pt", 2), ("epoch_1_model.pt", 3), ("epoch_2_model.pt", 4)):
        sd, tables[name] = _state_dict(seed, V=5000, D=64)
        torch.save(sd, models / name)
    res = load_embeddings_by_epoch(str(tmp_path), device="cuda", verbose=False)
    for seed in (7, 8):
        q = torch.randn(24, 64, generator=torch.Generator().manual_seed(seed)).to(torch.bfloat16).cuda()
        outs = res.scan_all_epochs(q, 10)
        assert list(outs) == ["initial", "epoch_0", "epoch_1", "epoch_2"]
        for name, out in outs.items():
            want = res.scan(name, q, 10)
            assert torch.equal(out.topk_idx, want.topk_idx) and torch.equal(out.topk_val, want.topk_val)
            torch.testing.assert_close(out.stats, want.stats, rtol=0, atol=0)
    assert len(res._graphs) == 1
