"""Streaming front end for host-resident query batches: overlaps the host->device copy of batch
i+1 and the device->host copy of result i-1 with the scan of batch i (two CUDA streams, two
device buffers, events -- no host synchronisation inside the loop except on the result that is
handed back, `lag` batches late).  This is the path a caller with queries in pinned host memory uses; `bench.py`
measures its `e2e` number through it.  With a sharded scanner every rank uploads 1/N of the batch and
the ranks all-gather it over NVLink."""
from __future__ import annotations

import collections
from typing import Iterable, Iterator, Optional, Tuple

import torch

from . import ops


class HostQueryPipeline:
    def __init__(self, table: torch.Tensor, k: int, *, normalize: bool = True, scale: float = 1.0,
                 inv_norm_t: Optional[torch.Tensor] = None, scanner=None, lag: int = 3,
                 reuse_host_buffers: bool = False):
        if not table.is_cuda:
            raise RuntimeError("table must be a CUDA tensor (no CPU fallback)")
        self.table, self.k, self.normalize, self.scale = table, int(k), normalize, float(scale)
        self.device = table.device
        self.scanner = scanner            # an optional sharded.ShardedConceptScan
        # results handed back `lag` batches late: the host thread may run that far ahead of the
        # device, which hides its jitter (8 ranks + NCCL proxy threads share the host cores)
        self.lag = max(1, int(lag))
        # False: every result is a fresh pinned tensor the caller owns.  True: results rotate through
        # lag + 2 pinned buffer sets allocated once -- a yielded result is then valid until lag + 1
        # more have been yielded (no pinned allocation in the loop: cudaHostAlloc maps the block into
        # every visible GPU and costs milliseconds on an 8-GPU box)
        self.reuse_host_buffers = bool(reuse_host_buffers)
        self._host_ring, self._host_next = [], 0
        self.inv_norm_t = inv_norm_t
        if normalize and inv_norm_t is None and scanner is None:
            self.inv_norm_t = ops.row_inv_norm(table)
        self.copy_stream = torch.cuda.Stream(self.device)
        self._bufs = [None, None]

    def _scan(self, q: torch.Tensor, labels) -> ops.ScanOutput:
        if self.scanner is not None:
            return self.scanner.scan(q, self.k, normalize_q=self.normalize, scale=self.scale, labels=labels)
        return ops.concept_scan(q, self.table, self.k, normalize_q=self.normalize,
                                normalize_t=self.normalize, scale=self.scale, labels=labels,
                                inv_norm_t=self.inv_norm_t)

    def run(self, host_batches: Iterable[torch.Tensor], labels: Optional[torch.Tensor] = None
            ) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        """Yields (topk_val, topk_idx, stats) as pinned HOST tensors, one per input batch, in
        order.  Each host batch should be pinned for the copies to be asynchronous."""
        main = torch.cuda.current_stream(self.device)
        pending = collections.deque()     # (host results, event) of the batches in flight
        it = iter(host_batches)
        nxt = next(it, None)
        slot = 0
        staged = None
        if nxt is not None:
            staged = self._stage(nxt, slot, main)
        while staged is not None:
            q_dev, ready, sliced = staged
            nxt = next(it, None)
            slot ^= 1
            staged = self._stage(nxt, slot, main) if nxt is not None else None   # H2D of i+1
            main.wait_event(ready)
            if sliced:                    # on the main stream: one communicator, one order of collectives
                self.scanner.all_gather_rows(q_dev)
            out = self._scan(q_dev, labels)                                      # scan of i
            done = torch.cuda.Event()
            done.record(main)
            self.copy_stream.wait_event(done)
            with torch.cuda.stream(self.copy_stream):                            # D2H of i
                host = tuple(h.copy_(t, non_blocking=True) for h, t in
                             zip(self._host_set(out), (out.topk_val, out.topk_idx, out.stats)))
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

    def _host_set(self, out):
        shapes = [(t.shape, t.dtype) for t in (out.topk_val, out.topk_idx, out.stats)]
        if not self.reuse_host_buffers:
            return [torch.empty(sh, dtype=dt, pin_memory=True) for sh, dt in shapes]
        if not self._host_ring or [(h.shape, h.dtype) for h in self._host_ring[0]] != shapes:
            self._host_ring = [[torch.empty(sh, dtype=dt, pin_memory=True) for sh, dt in shapes]
                               for _ in range(self.lag + 2)]
            self._host_next = 0
        hs = self._host_ring[self._host_next]
        self._host_next = (self._host_next + 1) % len(self._host_ring)
        return hs

    def _stage(self, host_q: torch.Tensor, slot: int, main):
        buf = self._bufs[slot]
        if buf is None or buf.shape != host_q.shape or buf.dtype != host_q.dtype:
            buf = torch.empty(host_q.shape, dtype=host_q.dtype, device=self.device)
            self._bufs[slot] = buf
        # the buffer may still be read by the scan two batches ago
        self.copy_stream.wait_stream(main)
        world = getattr(self.scanner, "world", 1)
        sliced = world > 1 and host_q.shape[0] % world == 0
        with torch.cuda.stream(self.copy_stream):
            if sliced:
                # Sharded scan: every rank needs the whole batch.  Each rank uploads only its 1/N
                # row slice over its own PCIe link; `run` all-gathers the slices over NVLink
                # right before the scan, instead of N full copies competing for host bandwidth.
                rows = host_q.shape[0] // world
                r = self.scanner.rank
                buf[r * rows:(r + 1) * rows].copy_(host_q[r * rows:(r + 1) * rows], non_blocking=True)
            else:
                buf.copy_(host_q, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return buf, ev, sliced
