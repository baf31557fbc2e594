"""GPU: the vocab-sharded path.  With one GPU only world_size=1 is exercised (scan -> merge of a
single record, no collective).  With >= 2 GPUs two ranks are spawned over NCCL and the merged
answer must equal the unsharded scan on the same inputs: indices bit-for-bit, LSE to 1e-6."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs(Q=304):
    g = torch.Generator().manual_seed(77)
    V, D, k = 6001, 256, 50                         # V odd: ragged shards; 304 = 8 * 38 rows
    q = torch.randn(Q, D, generator=g).to(torch.bfloat16)
    t = torch.randn(V, D, generator=g).to(torch.bfloat16)
    t[4000:4020] = t[:20]                           # exact ties across the shard boundary
    labels = torch.randint(0, V, (Q,), generator=g)
    labels[::9] = -100
    return q, t, labels, k


def _worker(rank, world, port, ret, exchange="auto"):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import multimodal_concept_learning_b200 as mcl
        from multimodal_concept_learning_b200.sharded import ShardedConceptScan, shard_rows
        q, t, labels, k = _inputs()
        qd, td, ld = q.cuda(), t.cuda(), labels.cuda()
        lo, hi = shard_rows(t.shape[0], world, rank)
        sc = ShardedConceptScan(td[lo:hi].contiguous(), t.shape[0], exchange=exchange)
        for _ in range(3):                           # reuse of the communicator and gather buffer
            out = sc.scan(qd, k, scale=20.0, labels=ld, label_smoothing=0.1)
        full = mcl.concept_scan(qd, td, k, scale=20.0, labels=ld, label_smoothing=0.1)
        torch.cuda.synchronize()
        assert torch.equal(out.topk_idx, full.topk_idx), "sharded indices differ"
        assert torch.equal(out.topk_val, full.topk_val), "sharded values differ"
        torch.testing.assert_close(out.lse, full.lse, rtol=1e-6, atol=1e-5)
        torch.testing.assert_close(out.loss, full.loss, rtol=1e-6, atol=1e-6)
        # a row count the world size does not divide: the all-gather merge instead of the row exchange
        q2, _, labels2, _ = _inputs(301)
        out2 = sc.scan(q2.cuda(), k, scale=20.0, labels=labels2.cuda())
        full2 = mcl.concept_scan(q2.cuda(), td, k, scale=20.0, labels=labels2.cuda())
        assert torch.equal(out2.topk_idx, full2.topk_idx) and torch.equal(out2.topk_val, full2.topk_val)
        torch.testing.assert_close(out2.lse, full2.lse, rtol=1e-6, atol=1e-5)
        # k = 1 (running-argmax epilogue) through the sharded path
        out3 = sc.scan(qd, 1, normalize_q=False, labels=ld)
        full3 = mcl.concept_scan(qd, td, 1, normalize_q=False, inv_norm_t=mcl.row_inv_norm(td), labels=ld)
        assert torch.equal(out3.topk_idx, full3.topk_idx)
        torch.testing.assert_close(out3.loss, full3.loss, rtol=1e-6, atol=1e-6)
        # host batches through the streaming pipeline: each rank uploads 1/world of the rows and
        # the library all-gathers them (Q = 304 rows divides by 2, 4 and 8)
        from multimodal_concept_learning_b200.pipeline import HostQueryPipeline
        pipe = HostQueryPipeline(td[lo:hi], k, scale=20.0, scanner=sc)
        qh = q.pin_memory()
        got = list(pipe.run([qh, qh.flip(0).contiguous().pin_memory(), qh]))
        assert len(got) == 3
        assert torch.equal(got[0][1], full.topk_idx.cpu()) and torch.equal(got[2][0], full.topk_val.cpu())
        assert torch.equal(got[1][1], full.topk_idx.cpu().flip(0)), "second (row-reversed) batch differs"
        pipe.close()
        # every rank keeps only the rows it merged: 1/world of the result per rank, no final all-gather;
        # more batches than staging slots, so the slot ring wraps
        pipe = HostQueryPipeline(td[lo:hi], k, scale=20.0, scanner=sc, local_rows=True, lag=2)
        r0, r1 = pipe.row_range(qh.shape[0])
        split = world > 2 or exchange != "nccl"          # the peer-memory exchange splits the rows at world 2 too
        assert (r0, r1) == ((rank * 304 // world, (rank + 1) * 304 // world) if split else (0, 304))
        flipped = qh.flip(0).contiguous().pin_memory()
        got = list(pipe.run([qh, flipped] * 6))
        assert len(got) == 12
        for i, (v, ix, st) in enumerate(got):
            want_idx = full.topk_idx.cpu() if i % 2 == 0 else full.topk_idx.cpu().flip(0)
            want_val = full.topk_val.cpu() if i % 2 == 0 else full.topk_val.cpu().flip(0)
            assert torch.equal(ix, want_idx[r0:r1]) and torch.equal(v, want_val[r0:r1]), f"batch {i}"
        pipe.close()
        # full scans after scans that kept only the local rows, on the same peer blocks: the two arrival
        # counters of the peer-memory exchange advance at different rates
        for _ in range(2):
            again = sc.scan(qd, k, scale=20.0, labels=ld, label_smoothing=0.1)
        part = sc.scan(qd, k, scale=20.0, labels=ld, local_rows_only=True)
        last = sc.scan(qd, k, scale=20.0, labels=ld, label_smoothing=0.1)
        torch.cuda.synchronize()
        assert torch.equal(again.topk_idx, full.topk_idx) and torch.equal(last.topk_val, full.topk_val)
        p0, p1 = sc.local_rows_for(304, k)
        assert torch.equal(part.topk_idx[p0:p1], full.topk_idx[p0:p1])
        sc.close()
        ret[rank] = "ok"
    except Exception as e:  # pragma: no cover
        ret[rank] = f"{type(e).__name__}: {e}"
    finally:
        dist.destroy_process_group()


def test_world1_sharded_equals_plain(lib_built):
    import multimodal_concept_learning_b200 as mcl
    from multimodal_concept_learning_b200.sharded import ShardedConceptScan
    q, t, labels, k = _inputs()
    sc = ShardedConceptScan(t.cuda(), t.shape[0])
    out = sc.scan(q.cuda(), k, scale=20.0, labels=labels.cuda())
    full = mcl.concept_scan(q.cuda(), t.cuda(), k, scale=20.0, labels=labels.cuda())
    assert torch.equal(out.topk_idx, full.topk_idx) and torch.equal(out.topk_val, full.topk_val)
    torch.testing.assert_close(out.stats, full.stats, rtol=1e-6, atol=1e-6)


# "auto": the result exchange over peer memory (mcl_concept_scan_sharded_p2p) wherever the shape allows
# it -- the 304-row scans with k = 50, and k = 1 at worlds 2 and 4 -- and NCCL elsewhere (301 rows; k = 1 at
# world 8); "nccl": 2 = all-gather merge; 4, 8 = row-exchange merge
@pytest.mark.parametrize("exchange", ["auto", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_rank_nccl_sharded_equals_unsharded(lib_built, world, exchange):
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs >= {world} GPUs")
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret, exchange), nprocs=world, join=True)
        assert dict(ret) == {r: "ok" for r in range(world)}
