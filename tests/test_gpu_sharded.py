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


def _inputs():
    g = torch.Generator().manual_seed(77)
    Q, V, D, k = 300, 6001, 256, 50                 # V odd: ragged shards
    q = torch.randn(Q, D, generator=g).to(torch.bfloat16)
    t = torch.randn(V, D, generator=g).to(torch.bfloat16)
    t[4000:4020] = t[:20]                           # exact ties across the shard boundary
    labels = torch.randint(0, V, (Q,), generator=g)
    labels[::9] = -100
    return q, t, labels, k


def _worker(rank, world, port, ret):
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
        sc = ShardedConceptScan(td[lo:hi].contiguous(), t.shape[0])
        for _ in range(3):                           # reuse of the communicator and gather buffer
            out = sc.scan(qd, k, scale=20.0, labels=ld, label_smoothing=0.1)
        full = mcl.concept_scan(qd, td, k, scale=20.0, labels=ld, label_smoothing=0.1)
        torch.cuda.synchronize()
        assert torch.equal(out.topk_idx, full.topk_idx), "sharded indices differ"
        assert torch.equal(out.topk_val, full.topk_val), "sharded values differ"
        torch.testing.assert_close(out.lse, full.lse, rtol=1e-6, atol=1e-5)
        torch.testing.assert_close(out.loss, full.loss, rtol=1e-6, atol=1e-6)
        # host batches through the streaming pipeline: each rank uploads 1/world of the rows and
        # the library all-gathers them (Q = 300 rows divides by 2 and 4)
        from multimodal_concept_learning_b200.pipeline import HostQueryPipeline
        pipe = HostQueryPipeline(td[lo:hi], k, scale=20.0, scanner=sc)
        qh = q.pin_memory()
        got = list(pipe.run([qh, qh.flip(0).contiguous().pin_memory(), qh]))
        assert len(got) == 3
        assert torch.equal(got[0][1], full.topk_idx.cpu()) and torch.equal(got[2][0], full.topk_val.cpu())
        assert torch.equal(got[1][1], full.topk_idx.cpu().flip(0)), "second (row-reversed) batch differs"
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


@pytest.mark.parametrize("world", [2, 4])      # 2: all-gather merge; 4: row-exchange merge
def test_multi_rank_nccl_sharded_equals_unsharded(lib_built, world):
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs >= {world} GPUs")
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {r: "ok" for r in range(world)}
