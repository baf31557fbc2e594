"""Vocabulary-row sharded concept scan over the GPUs of one NVSwitch box.

One process per GPU (``torch.distributed``; backend ``nccl`` on the box).  Rank r holds
table rows ``[r*ceil(V/N), min(V, (r+1)*ceil(V/N)))``; queries are replicated.  Every scan is
three operations enqueued on torch's current stream by ONE C call
(``mcl_concept_scan_sharded``): the local fused scan, a single ``ncclAllGather`` of the packed
per-rank record ``(k values, k indices, m, s, sum_z, z_label)`` per query -- Q*(12k+16) bytes
per rank, latency-bound on NVLink 5 -- and the merge kernel.  ``torch.distributed`` is used
for plumbing only: it carries the 128-byte ``ncclUniqueId`` of the library's private
communicator from rank 0 to the others (torch does not expose its own ``ncclComm_t``)."""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from . import ops
from ._lib import check, load

UNIQUE_ID_BYTES = 128


def shard_rows(V: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range of ``rank``: [lo, hi).  The last ranks may be short or empty."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad world/rank {world}/{rank}")
    per = -(-V // world)
    return min(V, rank * per), min(V, (rank + 1) * per)


def exchange_unique_id(make_id: Callable[[], bytes], group=None, device=None) -> bytes:
    """Rank 0 of ``group`` creates the communicator id, everybody receives it.  Works on any
    backend (the CPU tests drive it over gloo)."""
    rank = dist.get_rank(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) \
            if dist.get_backend(group) == "nccl" else torch.device("cpu")
    buf = torch.zeros(UNIQUE_ID_BYTES, dtype=torch.uint8, device=device)
    if rank == 0:
        raw = make_id()
        if len(raw) != UNIQUE_ID_BYTES:
            raise ValueError(f"unique id must be {UNIQUE_ID_BYTES} bytes")
        buf.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(buf, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return bytes(buf.cpu().tolist())


def _nccl_unique_id() -> bytes:
    raw = C.create_string_buffer(UNIQUE_ID_BYTES)
    check(load().mcl_comm_unique_id(raw))
    return raw.raw


class ShardedConceptScan:
    """Holds this rank's table shard, its cached inverse norms, the private NCCL communicator
    and the reusable gather buffer."""

    def __init__(self, table_shard: Tensor, vocab_total: int, *, normalize_t: bool = True,
                 group=None):
        if not table_shard.is_cuda:
            raise RuntimeError("table shard must be a CUDA tensor (no CPU fallback)")
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.vocab_total = int(vocab_total)
        self.lo, self.hi = shard_rows(self.vocab_total, self.world, self.rank)
        if table_shard.shape[0] != self.hi - self.lo:
            raise ValueError(f"rank {self.rank} expects rows [{self.lo},{self.hi}) "
                             f"= {self.hi - self.lo}, got {table_shard.shape[0]}")
        if self.hi - self.lo < 1:
            raise ValueError("empty table shard: use fewer ranks than table rows / k")
        self.table = ops._rowmajor(table_shard)
        self.device = table_shard.device
        self.inv_norm_t = ops.row_inv_norm(self.table) if normalize_t else None
        self._comm = C.c_void_p(None)
        self._gather = None
        if self.world > 1:
            uid = exchange_unique_id(_nccl_unique_id, group)
            with torch.cuda.device(self.device):
                check(load().mcl_comm_init(uid, self.world, self.rank, C.byref(self._comm)))

    def close(self):
        if self._comm:
            check(load().mcl_comm_destroy(self._comm))
            self._comm = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def all_gather_rows(self, buf: Tensor) -> None:
        """In-place all-gather of ``buf [world*rows, ...]`` on torch's current stream: rank r
        contributes rows ``[r*rows, (r+1)*rows)``.  Uses the library's own communicator, so it is
        ordered with the scans' collectives on the same stream."""
        if self.world == 1:
            return
        if not buf.is_contiguous() or buf.shape[0] % self.world:
            raise ValueError("buffer must be contiguous with a row count divisible by the world size")
        per = buf.numel() * buf.element_size() // self.world
        with torch.cuda.device(self.device):
            check(load().mcl_comm_all_gather(self._comm, buf.data_ptr() + self.rank * per, buf.data_ptr(), per,
                                             ops._stream(self.device)))

    def scan(self, q: Tensor, k: int, *, normalize_q: bool = True, scale: float = 1.0,
             labels: Optional[Tensor] = None, label_smoothing: float = 0.0,
             inv_norm_q: Optional[Tensor] = None) -> ops.ScanOutput:
        """Every rank passes the same ``q`` (and labels, GLOBAL row ids) and receives the same
        merged answer."""
        lib = load()
        dev = self.device
        if not q.is_cuda or q.dtype != self.table.dtype:
            raise TypeError("q must be a CUDA tensor of the table's dtype")
        kk = int(k)
        if not 1 <= kk <= min(ops.MCL_MAX_K, self.hi - self.lo):
            raise ValueError(f"k={k} must be in [1, min(rows per shard, {ops.MCL_MAX_K})]")
        q = ops._rowmajor(q)
        Q, D = q.shape
        if normalize_q and inv_norm_q is None:
            inv_norm_q = ops.row_inv_norm(q)
        if labels is not None:
            labels = labels.to(device=dev, dtype=torch.int64).contiguous()
        code = ops._dtype_code(q)
        val = torch.empty((Q, kk), dtype=torch.float32, device=dev)
        idx = torch.empty((Q, kk), dtype=torch.int64, device=dev)
        stats = torch.empty((Q, 4), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            gbytes = lib.mcl_sharded_gather_bytes(Q, kk, self.world)
            if self._gather is None or self._gather.numel() < gbytes:
                self._gather = torch.empty(gbytes, dtype=torch.uint8, device=dev)
            ws_bytes = lib.mcl_scan_workspace_bytes(Q, self.hi - self.lo, D, kk, code)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            check(lib.mcl_concept_scan_sharded(
                q.data_ptr(), self.table.data_ptr(), code, Q, self.hi - self.lo, D, q.stride(0),
                self.table.stride(0), ops._ptr(inv_norm_q), ops._ptr(self.inv_norm_t), float(scale),
                kk, self.lo, ops._ptr(labels), val.data_ptr(), idx.data_ptr(), stats.data_ptr(),
                ws.data_ptr(), ws_bytes, self._gather.data_ptr(), gbytes, self._comm, self.world,
                self.rank, ops._stream(dev)))
        return ops.ScanOutput(val, idx, stats, self.vocab_total, labels, float(label_smoothing))
