"""Vocabulary-row sharded concept scan over the GPUs of one NVSwitch box.

One process per GPU (``torch.distributed``; backend ``nccl`` on the box).  Rank r holds
table rows ``[r*ceil(V/N), min(V, (r+1)*ceil(V/N)))``; queries are replicated.  Every scan is
ONE C call that enqueues, on torch's current stream, the local fused scan, the exchange of the
per-rank records ``(k values, k indices, m, s, sum_z, z_label)`` per query -- rank j collects and
merges query rows ``[j*Q/N, (j+1)*Q/N)`` -- and the distribution of the merged rows:

* ``mcl_concept_scan_sharded_p2p`` (default): stores into the peers' memory over NVLink, arrival
  counters, stream waits; no NCCL call in the step (``csrc/p2p_exchange.cu``);
* ``mcl_concept_scan_sharded_ex`` (``exchange="nccl"``, and whenever the shape rules the peer-memory
  path out): one ``ncclAllGather`` + full merge at world 2, grouped send/recv + all-gather above.

``torch.distributed`` is plumbing only: it carries the 128-byte ``ncclUniqueId`` of the library's
private communicator and the CUDA IPC handles of the peer-memory blocks between the ranks."""
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
    """Holds this rank's table shard, its cached inverse norms, the private NCCL communicator,
    the peer-memory blocks of the result exchange and the reusable gather buffer."""

    MAX_EPOCH = 1 << 31               # scans per set of peer-memory blocks (include/mcl.h)

    def __init__(self, table_shard: Tensor, vocab_total: int, *, normalize_t: bool = True,
                 group=None, exchange: str = "auto"):
        """``exchange``: how the per-rank results meet -- ``"p2p"`` stores them straight into the
        peers' memory over NVLink (``mcl_concept_scan_sharded_p2p``: two small kernels and two
        stream waits per scan), ``"nccl"`` uses the library's communicator (one all-gather, or a
        grouped send/recv + all-gather for world > 2), ``"auto"`` (default) takes the peer-memory
        path whenever the shape allows it (Q % world == 0, (Q / world) * k % 4 == 0)."""
        if exchange not in ("auto", "p2p", "nccl"):
            raise ValueError("exchange must be 'auto', 'p2p' or 'nccl'")
        self.exchange = exchange
        self._boards = {}                 # (Q, k) -> peer-memory blocks of the p2p exchange
        if not table_shard.is_cuda:
            raise RuntimeError("table shard must be a CUDA tensor (no CPU fallback)")
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.vocab_total = int(vocab_total)
        self.lo, self.hi = shard_rows(self.vocab_total, self.world, self.rank)
        # Rank-independent checks first: an exception on ONE rank before a collective would leave
        # the others blocked in it.  The shortest shard is the last one.
        self.min_rows = min(h - l for l, h in (shard_rows(self.vocab_total, self.world, r)
                                               for r in range(self.world)))
        if self.min_rows < 1:
            raise ValueError(f"{self.vocab_total} table rows leave a rank of {self.world} without rows: "
                             "use fewer ranks")
        if self.vocab_total > (1 << 32):
            raise ValueError("the rank merge packs global table rows into 32 bits: vocab_total <= 2^32")
        if table_shard.shape[0] != self.hi - self.lo:
            raise ValueError(f"rank {self.rank} expects rows [{self.lo},{self.hi}) "
                             f"= {self.hi - self.lo}, got {table_shard.shape[0]}")
        self.table = ops._rowmajor(table_shard)
        self.device = table_shard.device
        self.inv_norm_t = ops.row_inv_norm(self.table) if normalize_t else None
        self._comm = C.c_void_p(None)
        self._gather = None
        if self.world > 1:
            uid = exchange_unique_id(_nccl_unique_id, group)
            with torch.cuda.device(self.device):
                check(load().mcl_comm_init(uid, self.world, self.rank, C.byref(self._comm)))

    def _free_boards(self, boards) -> None:
        """Collective: every rank frees the same boards.  Order required by CUDA IPC: all stores
        have landed, every importer has closed its mapping, only then does the owner free."""
        lib = load()
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)              # nobody stores into a block that is being freed
        with torch.cuda.device(self.device):
            for board in boards:
                for r, p in enumerate(board["ptrs"]):
                    if r != self.rank and p:
                        lib.mcl_peer_close(p)
        dist.barrier(group=self.group)              # no mapping of a block is left when its owner frees it
        with torch.cuda.device(self.device):
            for board in boards:
                lib.mcl_peer_free(board["base"])

    def close(self):
        """Frees the peer-memory blocks and the communicator.  A COLLECTIVE when the peer-memory
        exchange was used: call it on every rank (garbage collection does not do it for you -- a
        finaliser must not enter a barrier)."""
        if getattr(self, "_boards", None):
            boards, self._boards = list(self._boards.values()), {}
            self._free_boards(boards)
        if self._comm:
            check(load().mcl_comm_destroy(self._comm))
            self._comm = C.c_void_p(None)

    def _p2p_board(self, Q: int, k: int):
        """The peer-memory blocks of the result exchange for one (Q, k) shape, or None when the
        shape (or the ``exchange`` setting) rules the path out.  Creating a board is a collective:
        every rank reaches it in the same scan."""
        if self.world < 2 or self.exchange == "nccl":
            return None
        key = (int(Q), int(k))
        board = self._boards.get(key)
        if board is not None and board["epoch"] >= self.MAX_EPOCH:
            # the library counts the scans on a set of blocks in 32 bits: start over on fresh blocks
            # (every rank gets here in the same scan: the epochs advance in lockstep)
            self._free_boards([self._boards.pop(key)])
            board = None
        if board is not None:
            return board
        lib = load()
        nbytes = int(lib.mcl_sharded_p2p_block_bytes(Q, k, self.world))
        if nbytes == 0:
            if self.exchange == "p2p":
                raise ValueError(f"the peer-memory exchange needs Q % world == 0 and (Q / world) * k % 4 == 0 "
                                 f"(Q={Q}, k={k}, world={self.world})")
            return None
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        with torch.cuda.device(self.device):
            check(lib.mcl_peer_alloc(nbytes, C.byref(ptr), handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=self.group)
        ptrs = [None] * self.world
        with torch.cuda.device(self.device):
            for r, h in enumerate(handles):
                if r == self.rank:
                    ptrs[r] = ptr.value
                else:
                    pp = C.c_void_p()
                    check(lib.mcl_peer_open(h, C.byref(pp)))
                    ptrs[r] = pp.value
        dist.barrier(group=self.group)              # every mapping exists before the first store
        board = {"base": ptr.value, "ptrs": ptrs, "bytes": nbytes, "epoch": 0, "full_epoch": 0,
                 "array": (C.c_void_p * self.world)(*ptrs)}
        self._boards[key] = board
        return board

    def __del__(self):
        # never a collective from a finaliser (the ranks collect garbage at different times): the
        # peer-memory blocks of an unclosed scanner are left to process exit
        try:
            if getattr(self, "_boards", None):
                import warnings
                warnings.warn("ShardedConceptScan dropped without close(): its peer-memory blocks stay "
                              "allocated until the process exits", ResourceWarning)
                return
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

    def local_rows(self, Q: int) -> Tuple[int, int]:
        """The query rows this rank merges itself under the row exchange (world > 2 and
        Q % world == 0); every row otherwise."""
        if self.world > 2 and Q % self.world == 0:
            per = Q // self.world
            return self.rank * per, (self.rank + 1) * per
        return 0, Q

    def local_rows_for(self, Q: int, k: int) -> Tuple[int, int]:
        """As :meth:`local_rows`, for a scan with this ``k``: the peer-memory exchange splits the rows
        over the ranks for every world >= 2."""
        if self.world >= 2 and self.exchange != "nccl" and \
                int(load().mcl_sharded_p2p_block_bytes(Q, k, self.world)) > 0:
            per = Q // self.world
            return self.rank * per, (self.rank + 1) * per
        return self.local_rows(Q)

    def scan(self, q: Tensor, k: int, *, normalize_q: bool = True, scale: float = 1.0,
             labels: Optional[Tensor] = None, label_smoothing: float = 0.0,
             inv_norm_q: Optional[Tensor] = None, local_rows_only: bool = False) -> ops.ScanOutput:
        """Every rank passes the same ``q`` (and labels, GLOBAL row ids) and receives the same
        merged answer.  ``local_rows_only``: only the rows of :meth:`local_rows` are valid in the
        outputs (the final all-gather of the merged rows is skipped) -- for consumers that take
        each rank's row range separately, e.g. one host copy of 1/world of the result per rank."""
        lib = load()
        dev = self.device
        if not q.is_cuda or q.dtype != self.table.dtype:
            raise TypeError("q must be a CUDA tensor of the table's dtype")
        kk = int(k)
        if not 1 <= kk <= min(ops.MCL_MAX_K, self.min_rows):       # the same verdict on every rank
            raise ValueError(f"k={k} must be in [1, min(rows of the shortest shard = {self.min_rows}, "
                             f"{ops.MCL_MAX_K})]")
        q = ops._rowmajor(q)
        Q, D = q.shape
        flags = (1 if local_rows_only else 0) | (2 if (normalize_q and inv_norm_q is None) else 0)
        if labels is not None:
            labels = labels.to(device=dev, dtype=torch.int64).contiguous()
        code = ops._dtype_code(q)
        val = torch.empty((Q, kk), dtype=torch.float32, device=dev)
        idx = torch.empty((Q, kk), dtype=torch.int64, device=dev)
        stats = torch.empty((Q, 4), dtype=torch.float32, device=dev)
        board = self._p2p_board(Q, kk)
        if board is not None:
            board["epoch"] += 1
            if not local_rows_only:                  # the merged rows travel (and are counted) only then
                board["full_epoch"] += 1
            with torch.cuda.device(dev):
                ws_bytes = lib.mcl_scan_workspace_bytes(Q, self.hi - self.lo, D, kk, code)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                check(lib.mcl_concept_scan_sharded_p2p(
                    q.data_ptr(), self.table.data_ptr(), code, Q, self.hi - self.lo, D, q.stride(0),
                    self.table.stride(0), ops._ptr(inv_norm_q), ops._ptr(self.inv_norm_t), float(scale),
                    kk, self.lo, ops._ptr(labels), val.data_ptr(), idx.data_ptr(), stats.data_ptr(),
                    ws.data_ptr(), ws_bytes, board["array"], board["bytes"], self.world, self.rank,
                    board["epoch"], board["full_epoch"], flags, ops._stream(dev)))
            return ops.ScanOutput(val, idx, stats, self.vocab_total, labels, float(label_smoothing))
        with torch.cuda.device(dev):
            gbytes = lib.mcl_sharded_gather_bytes(Q, kk, self.world)
            if self._gather is None or self._gather.numel() < gbytes:
                self._gather = torch.empty(gbytes, dtype=torch.uint8, device=dev)
            ws_bytes = lib.mcl_scan_workspace_bytes(Q, self.hi - self.lo, D, kk, code)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            check(lib.mcl_concept_scan_sharded_ex(
                q.data_ptr(), self.table.data_ptr(), code, Q, self.hi - self.lo, D, q.stride(0),
                self.table.stride(0), ops._ptr(inv_norm_q), ops._ptr(self.inv_norm_t), float(scale),
                kk, self.lo, ops._ptr(labels), val.data_ptr(), idx.data_ptr(), stats.data_ptr(),
                ws.data_ptr(), ws_bytes, self._gather.data_ptr(), gbytes, self._comm, self.world,
                self.rank, flags, ops._stream(dev)))
        return ops.ScanOutput(val, idx, stats, self.vocab_total, labels, float(label_smoothing))


class _RawCudaBuffer:
    """``__cuda_array_interface__`` view of a raw device pointer (a library-owned or IPC-mapped
    block), so that torch can alias it without owning it."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                         "version": 2}


class QueryBoard:
    """Replicates host-resident query batches on every rank of a sharded scan WITHOUT SMs and
    without touching the scan's stream order: each rank uploads 1/N of the batch over its own PCIe
    link into a ring of staging slots and pushes that slice into the same slot of every peer with
    copy-engine copies over NVLink (CUDA IPC mappings, ``mcl_peer_*``), then writes a step counter
    into the peer's flag word; the consumer's stream waits for its flag words with a stream memory
    operation (``mcl_stream_wait_value32``).  All of it runs on a side stream while the previous
    batch is being scanned.

    Slot reuse needs no back-pressure: every sharded scan ends in a collective over all ranks, so a
    rank that is enqueueing step n has seen every rank's GPU finish step n - lag - 1 (it holds that
    step's result on the host); ``slots >= lag + 2`` therefore never overwrites a batch that a peer
    still reads."""

    def __init__(self, scanner: "ShardedConceptScan", Q: int, D: int, dtype: torch.dtype, slots: int):
        if scanner.world < 2 or Q % scanner.world:
            raise ValueError("QueryBoard needs world > 1 and a row count the world size divides")
        lib = load()
        self.scanner, self.world, self.rank = scanner, scanner.world, scanner.rank
        self.Q, self.D, self.dtype, self.slots = int(Q), int(D), dtype, int(slots)
        self.device = scanner.device
        es = torch.empty(0, dtype=dtype).element_size()
        self.rows = self.Q // self.world
        self.slice_bytes = self.rows * self.D * es
        self.slot_bytes = self.Q * self.D * es
        self.flag_off = (self.slots * self.slot_bytes + 255) & ~255
        self.nbytes = self.flag_off + self.slots * self.world * 4
        self.step = 0
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        with torch.cuda.device(self.device):
            check(lib.mcl_peer_alloc(self.nbytes, C.byref(ptr), handle))
        self.base = ptr.value
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=scanner.group)
        self.peer_base = [None] * self.world
        with torch.cuda.device(self.device):
            for r, h in enumerate(handles):
                if r == self.rank:
                    self.peer_base[r] = self.base
                else:
                    pp = C.c_void_p()
                    check(lib.mcl_peer_open(h, C.byref(pp)))
                    self.peer_base[r] = pp.value
        self._holder = _RawCudaBuffer(self.base, self.nbytes)
        raw = torch.as_tensor(self._holder, device=self.device)
        self.batches = [raw[s * self.slot_bytes:(s + 1) * self.slot_bytes].view(dtype).view(self.Q, self.D)
                        for s in range(self.slots)]
        # step counters the flag copies read (pinned host ring, rewritten long after its copy ran)
        self.vals = torch.zeros(256, dtype=torch.int32).pin_memory()
        dist.barrier(group=scanner.group)           # every mapping exists before the first push

    def publish(self, host_q: Tensor, copy_stream: torch.cuda.Stream):
        """Enqueue on ``copy_stream``: upload this rank's row slice of ``host_q`` and push it (and the
        step counter) to every peer.  Returns (step, event after the local upload)."""
        lib = load()
        n = self.step
        self.step += 1
        slot = n % self.slots
        lo = self.rank * self.rows
        mine = self.batches[slot][lo:lo + self.rows]
        cs = copy_stream.cuda_stream
        with torch.cuda.device(self.device), torch.cuda.stream(copy_stream):
            mine.copy_(host_q[lo:lo + self.rows], non_blocking=True)
            up = torch.cuda.Event()
            up.record(copy_stream)
            self.vals[n % 256] = n + 1
            off = slot * self.slot_bytes + lo * self.D * mine.element_size()
            for r in range(self.world):
                if r == self.rank:
                    continue
                check(lib.mcl_memcpy_async(self.peer_base[r] + off, self.base + off, self.slice_bytes, cs))
            for r in range(self.world):
                if r == self.rank:
                    continue
                flag = self.peer_base[r] + self.flag_off + 4 * (slot * self.world + self.rank)
                check(lib.mcl_memcpy_async(flag, self.vals.data_ptr() + 4 * (n % 256), 4, cs))
        return n, up

    def wait(self, n: int, uploaded: torch.cuda.Event, stream: torch.cuda.Stream) -> Tensor:
        """Make ``stream`` wait until every rank's slice of step ``n`` has landed; returns the batch."""
        lib = load()
        slot = n % self.slots
        stream.wait_event(uploaded)
        with torch.cuda.device(self.device):
            for r in range(self.world):
                if r != self.rank:
                    check(lib.mcl_stream_wait_value32(stream.cuda_stream,
                                                      self.base + self.flag_off + 4 * (slot * self.world + r), n + 1))
        return self.batches[slot]

    def close(self):
        lib = load()
        if self.base is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.scanner.group)      # nobody pushes into a block that is being freed
        with torch.cuda.device(self.device):
            for r, p in enumerate(self.peer_base):
                if r != self.rank and p:
                    lib.mcl_peer_close(p)
        dist.barrier(group=self.scanner.group)      # every importer closed its mapping before the owner frees
        with torch.cuda.device(self.device):
            self.batches = []
            lib.mcl_peer_free(self.base)
        self.base = None
