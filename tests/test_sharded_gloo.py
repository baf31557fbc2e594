"""CPU, world_size 2 over gloo: the host side of the vocab-sharded path -- shard bounds, the
communicator-id exchange (`sharded.exchange_unique_id`) and the gather -> merge data flow
(record order = rank order), with the oracle standing in for the per-rank CUDA scan."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodal_concept_learning_b200.sharded import UNIQUE_ID_BYTES, exchange_unique_id, shard_rows
from oracle import concept_scan_ref as R


def test_shard_rows_partition():
    for V, world in [(152064, 8), (128256, 8), (50257, 3), (10, 4), (7, 8)]:
        spans = [shard_rows(V, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == V
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert spans == R.shard_bounds(V, world)
    assert shard_rows(152064, 8, 3) == (57024, 76032)       # 19008 rows per rank
    with pytest.raises(ValueError):
        shard_rows(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. communicator id: made on rank 0 only, identical everywhere afterwards
        uid = exchange_unique_id(lambda: bytes((7 * i + 3) % 256 for i in range(UNIQUE_ID_BYTES)))
        assert uid == bytes((7 * i + 3) % 256 for i in range(UNIQUE_ID_BYTES))
        # 2. per-rank scan of this rank's rows (oracle in place of the CUDA kernel), gather, merge
        g = torch.Generator().manual_seed(99)
        Q, V, D, k = 21, 333, 40, 9
        q, t = torch.randn(Q, D, generator=g), torch.randn(V, D, generator=g)
        t[200:205] = t[:5]                                  # ties across the shard boundary
        labels = torch.randint(0, V, (Q,), generator=g)
        labels[2] = -100
        lo, hi = shard_rows(V, world, rank)
        part = R.concept_scan_ref(q, t[lo:hi], k, index_base=lo, labels=labels, vocab_total=V, scale=5.0)
        rec = torch.cat([part.topk_val, part.topk_idx.double(),
                         torch.stack([part.m, part.s, part.sum_z, part.z_label], 1)], 1)
        gathered = [torch.empty_like(rec) for _ in range(world)]
        dist.all_gather(gathered, rec)
        vals = [x[:, :k] for x in gathered]
        idxs = [x[:, k:2 * k].long() for x in gathered]
        st = [x[:, 2 * k:] for x in gathered]
        val, idx, m, s, sz, zl = R.merge_ref(vals, idxs, [x[:, 0] for x in st], [x[:, 1] for x in st],
                                             [x[:, 2] for x in st], [x[:, 3] for x in st], k)
        full = R.concept_scan_ref(q, t, k, labels=labels, scale=5.0)
        assert torch.equal(idx, full.topk_idx)
        torch.testing.assert_close(val, full.topk_val, rtol=1e-12, atol=1e-12)
        torch.testing.assert_close(m + torch.log(s), full.lse, rtol=1e-12, atol=1e-12)
        torch.testing.assert_close(zl, full.z_label, rtol=1e-12, atol=1e-12)
        ret[rank] = "ok"
    except Exception as e:  # pragma: no cover
        ret[rank] = f"{type(e).__name__}: {e}"
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_exchange_and_merge():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: "ok", 1: "ok"}
