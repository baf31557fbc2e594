"""Streaming front end for host-resident query batches: overlaps the host->device copy of batch
i+1 and the device->host copy of result i-1 with the scan of batch i (two CUDA streams, a ring
of device buffers, events -- no host synchronisation inside the loop except on the result that
is handed back, `lag` batches late).  This is the path a caller with queries in pinned host
memory uses; `bench.py` measures its `e2e` number through it.

With a sharded scanner every rank uploads 1/N of the batch over its own PCIe link and pushes its
slice to the peers with copy-engine copies over NVLink (`sharded.QueryBoard`): no SM and no
collective on the scan's stream is spent on distributing the queries, and the copies of batch
i+1 run while batch i is scanned.  `local_rows=True` additionally returns, on every rank, only
the rows that rank merged (1/N of the result: no all-gather of the merged rows, 1/N of the
device->host bytes per rank)."""
from __future__ import annotations

import collections
from typing import Iterable, Iterator, Optional, Tuple

import torch

from . import ops


class HostQueryPipeline:
    def __init__(self, table: torch.Tensor, k: int, *, normalize: bool = True, scale: float = 1.0,
                 inv_norm_t: Optional[torch.Tensor] = None, scanner=None, lag: int = 3,
                 reuse_host_buffers: bool = False, local_rows: bool = False):
        if not table.is_cuda:
            raise RuntimeError("table must be a CUDA tensor (no CPU fallback)")
        self.table, self.k, self.normalize, self.scale = table, int(k), normalize, float(scale)
        self.device = table.device
        self.scanner = scanner            # an optional sharded.ShardedConceptScan
        self.world = getattr(scanner, "world", 1)
        # results handed back `lag` batches late: the host thread may run that far ahead of the
        # device, which hides its jitter (8 ranks + NCCL proxy threads share the host cores)
        self.lag = max(1, int(lag))
        # False: every result is a fresh pinned tensor the caller owns.  True: results rotate through
        # 2*lag + 2 pinned buffer sets allocated once -- a yielded result is then valid until lag + 1
        # more have been yielded: `lag` sets are always in flight, and the copies enqueued while the
        # caller holds a result for lag + 1 further yields touch lag + 1 more sets (no pinned
        # allocation in the loop: cudaHostAlloc maps the block into every visible GPU and costs
        # milliseconds on an 8-GPU box)
        self.reuse_host_buffers = bool(reuse_host_buffers)
        self.local_rows = bool(local_rows) and self.world > 1
        self._host_ring, self._host_next = [], 0
        self.inv_norm_t = inv_norm_t
        if normalize and inv_norm_t is None and scanner is None:
            self.inv_norm_t = ops.row_inv_norm(table)
        self.copy_stream = torch.cuda.Stream(self.device)
        self._bufs = [None] * (self.lag + 3)
        self._board = None

    # ---- introspection for bench.py ---------------------------------------------------------
    def _sliced(self, Q: int) -> bool:
        return self.world > 1 and Q % self.world == 0

    def h2d_bytes_per_step(self, host_q: torch.Tensor) -> int:
        """Host->device bytes of one step summed over all ranks: the batch once when the ranks
        upload slices, `world` times when every rank uploads all of it."""
        b = host_q.numel() * host_q.element_size()
        return b if (self.world == 1 or self._sliced(host_q.shape[0])) else b * self.world

    def describe(self) -> str:
        if self.world == 1:
            return "H2D of the batch and D2H of (val, idx, stats) overlapped with the scan on a copy stream"
        return ("each rank uploads 1/N of the batch and pushes it to the peers over NVLink with copy engines "
                "(QueryBoard); " + ("each rank returns the 1/N of the result rows it merged"
                                    if self.local_rows else "every rank returns the full result"))

    def row_range(self, Q: int) -> Tuple[int, int]:
        """Rows of the batch that the tensors yielded by :meth:`run` hold on this rank."""
        return self.scanner.local_rows_for(Q, self.k) if self.local_rows else (0, Q)

    def close(self):
        if self._board is not None:
            self._board.close()
            self._board = None

    # ---- the loop -----------------------------------------------------------------------------
    def _scan(self, q: torch.Tensor, labels) -> ops.ScanOutput:
        if self.scanner is not None:
            return self.scanner.scan(q, self.k, normalize_q=self.normalize, scale=self.scale, labels=labels,
                                     local_rows_only=self.local_rows)
        return ops.concept_scan(q, self.table, self.k, normalize_q=self.normalize,
                                normalize_t=self.normalize, scale=self.scale, labels=labels,
                                inv_norm_t=self.inv_norm_t)

    def run(self, host_batches: Iterable[torch.Tensor], labels: Optional[torch.Tensor] = None
            ) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        """Yields (topk_val, topk_idx, stats) as pinned HOST tensors, one per input batch, in
        order (rows :meth:`row_range` of the batch).  Each host batch should be pinned for the
        copies to be asynchronous.  With a sharded scanner every rank must iterate the same
        sequence of batch shapes."""
        main = torch.cuda.current_stream(self.device)
        pending = collections.deque()     # (host results, event) of the batches in flight
        it = iter(host_batches)
        nxt = next(it, None)
        slot = 0
        staged = None
        if nxt is not None:
            staged = self._stage(nxt, slot, main)
        while staged is not None:
            cur = staged
            nxt = next(it, None)
            slot = (slot + 1) % len(self._bufs)
            staged = self._stage(nxt, slot, main) if nxt is not None else None   # H2D (+ pushes) of i+1
            q_dev = cur()                                                        # main waits for batch i
            out = self._scan(q_dev, labels)                                      # scan of i
            r0, r1 = self.row_range(q_dev.shape[0])
            parts = (out.topk_val[r0:r1], out.topk_idx[r0:r1], out.stats[r0:r1])
            done = torch.cuda.Event()
            done.record(main)
            self.copy_stream.wait_event(done)
            with torch.cuda.stream(self.copy_stream):                            # D2H of i
                host = tuple(h.copy_(t, non_blocking=True) for h, t in zip(self._host_set(parts), parts))
                copied = torch.cuda.Event()
                copied.record(self.copy_stream)
            for t in (out.topk_val, out.topk_idx, out.stats):
                t.record_stream(self.copy_stream)
            pending.append((host, copied))
            while len(pending) > self.lag:
                res, ev = pending.popleft()
                ev.synchronize()
                yield res
        while pending:
            res, ev = pending.popleft()
            ev.synchronize()
            yield res

    def _host_set(self, parts):
        shapes = [(t.shape, t.dtype) for t in parts]
        if not self.reuse_host_buffers:
            return [torch.empty(sh, dtype=dt, pin_memory=True) for sh, dt in shapes]
        if not self._host_ring or [(h.shape, h.dtype) for h in self._host_ring[0]] != shapes:
            self._host_ring = [[torch.empty(sh, dtype=dt, pin_memory=True) for sh, dt in shapes]
                               for _ in range(2 * self.lag + 2)]
            self._host_next = 0
        hs = self._host_ring[self._host_next]
        self._host_next = (self._host_next + 1) % len(self._host_ring)
        return hs

    def _stage(self, host_q: torch.Tensor, slot: int, main):
        """Starts the upload of a batch on the copy stream; returns a function that makes `main`
        wait for it and hands back the device batch."""
        if self._sliced(host_q.shape[0]):
            from .sharded import QueryBoard
            b = self._board
            if b is None or (b.Q, b.D, b.dtype) != (host_q.shape[0], host_q.shape[1], host_q.dtype):
                if b is not None:
                    b.close()
                b = self._board = QueryBoard(self.scanner, host_q.shape[0], host_q.shape[1], host_q.dtype,
                                             slots=self.lag + 3)
            # (slot reuse is safe without waiting: see QueryBoard)
            n, up = b.publish(host_q, self.copy_stream)
            return lambda: b.wait(n, up, main)
        buf = self._bufs[slot]
        if buf is None or buf.shape != host_q.shape or buf.dtype != host_q.dtype:
            buf = torch.empty(host_q.shape, dtype=host_q.dtype, device=self.device)
            self._bufs[slot] = buf
        # the buffer may still be read by the scan lag + 3 batches ago
        self.copy_stream.wait_stream(main)
        with torch.cuda.stream(self.copy_stream):
            buf.copy_(host_q, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)

        def ready():
            main.wait_event(ev)
            return buf
        return ready
