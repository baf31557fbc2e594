"""Host-side operators over CUDA tensors: thin wrappers that allocate outputs/workspace with
torch, pass raw pointers to libmcl_sm100.so on torch's current stream, and register the
calls as torch custom ops (``torch.ops.mcl.*``) so they compose with the rest of a torch
program.  PyTorch is plumbing here (device memory, streams); all arithmetic of the path is
in the CUDA library.  CPU tensors are rejected -- there is no fallback."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import MCL_DTYPE_BF16, MCL_DTYPE_F32, MCL_MAX_K, check, load

IGNORE_INDEX = -100


def _dtype_code(t: Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return MCL_DTYPE_BF16
    if t.dtype == torch.float32:
        return MCL_DTYPE_F32
    raise TypeError(f"libmcl_sm100 takes bfloat16 (tcgen05 path) or float32 (check path), got {t.dtype}")


def _require_cuda(*ts: Optional[Tensor]) -> torch.device:
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("libmcl_sm100 operators need CUDA tensors (no CPU fallback); "
                               "move the tensor to the GPU first")
        if dev is not None and t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} vs {t.device}")
        dev = t.device
    assert dev is not None
    return dev


def _rowmajor(t: Tensor) -> Tensor:
    """Inner dim contiguous, 16-byte aligned base and pitch -- else a contiguous copy."""
    if t.dim() != 2:
        raise ValueError(f"expected a 2-D tensor, got {tuple(t.shape)}")
    es = t.element_size()
    ok = (t.stride(1) == 1 or t.shape[1] == 1) and t.stride(0) >= t.shape[1] \
        and (t.stride(0) * es) % 16 == 0 and t.data_ptr() % 16 == 0
    if ok:
        return t
    c = t.contiguous()
    if (c.stride(0) * es) % 16 != 0:   # pad the pitch to 16 bytes
        pad = (-c.shape[1]) % (16 // es)
        buf = torch.zeros((c.shape[0], c.shape[1] + pad), dtype=c.dtype, device=c.device)
        buf[:, :c.shape[1]] = c
        return buf[:, :c.shape[1]]
    return c


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# --------------------------------------------------------------------------------------
# custom ops
# --------------------------------------------------------------------------------------

@torch.library.custom_op("mcl::row_inv_norm", mutates_args=(), device_types="cuda")
def _row_inv_norm_op(x: Tensor) -> Tensor:
    x = _rowmajor(x)
    out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(load().mcl_row_inv_norm(x.data_ptr(), _dtype_code(x), x.shape[0], x.shape[1],
                                      x.stride(0), out.data_ptr(), _stream(x.device)))
    return out


@_row_inv_norm_op.register_fake
def _(x):
    return x.new_empty(x.shape[0], dtype=torch.float32)


@torch.library.custom_op("mcl::gather_mean", mutates_args=(), device_types="cuda")
def _gather_mean_op(table: Tensor, offsets: Tensor, ids: Tensor, normalize: bool) -> Tensor:
    table = _rowmajor(table)
    Q = offsets.numel() - 1
    D = table.shape[1]
    es = table.element_size()
    ld_out = D + ((-D) % (16 // es))
    buf = torch.empty((Q, ld_out), dtype=table.dtype, device=table.device)
    flag = torch.zeros(1, dtype=torch.int32, device=table.device)   # ids were range-checked by gather_mean()
    with torch.cuda.device(table.device):
        check(load().mcl_gather_mean(table.data_ptr(), _dtype_code(table), table.shape[0], D,
                                     table.stride(0), offsets.data_ptr(), ids.data_ptr(), Q,
                                     int(normalize), buf.data_ptr(), ld_out, flag.data_ptr(),
                                     _stream(table.device)))
    out = buf[:, :D]
    return out if ld_out == D else out.contiguous()


@_gather_mean_op.register_fake
def _(table, offsets, ids, normalize):
    return table.new_empty((offsets.numel() - 1, table.shape[1]))


@torch.library.custom_op("mcl::ce_from_stats", mutates_args=(), device_types="cuda")
def _ce_from_stats_op(stats: Tensor, labels: Tensor, label_smoothing: float, vocab: int) -> Tuple[Tensor, Tensor]:
    Q = stats.shape[0]
    rows = torch.empty(Q, dtype=torch.float32, device=stats.device)
    mean = torch.empty(2, dtype=torch.float32, device=stats.device)
    with torch.cuda.device(stats.device):
        check(load().mcl_ce_from_stats(stats.data_ptr(), labels.data_ptr(), Q, float(label_smoothing),
                                       int(vocab), rows.data_ptr(), mean.data_ptr(), _stream(stats.device)))
    return rows, mean


@_ce_from_stats_op.register_fake
def _(stats, labels, label_smoothing, vocab):
    return stats.new_empty(stats.shape[0]), stats.new_empty(2)


_WS_BYTES: dict = {}     # (device index, Q, V, D, k, dtype code) -> workspace bytes (a pure function of these)


def _concept_scan_impl(q: Tensor, table: Tensor, inv_norm_q: Optional[Tensor],
                       inv_norm_t: Optional[Tensor], labels: Optional[Tensor], scale: float,
                       k: int, index_base: int, softcap: float = 0.0,
                       normalize_q: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
    lib = load()
    dev = q.device
    Q, D = q.shape
    V = table.shape[0]
    code = _dtype_code(q)
    val = torch.empty((Q, k), dtype=torch.float32, device=dev)
    idx = torch.empty((Q, k), dtype=torch.int64, device=dev)
    stats = torch.empty((Q, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        key = (dev.index, Q, V, D, k, code)
        ws_bytes = _WS_BYTES.get(key)
        if ws_bytes is None:
            ws_bytes = _WS_BYTES[key] = lib.mcl_scan_workspace_bytes(Q, V, D, k, code)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        # normalize_q without caller-supplied norms: the library forms them (inside the scan kernel
        # for small batches -- MCL_SCAN_NORMALIZE_Q)
        flags = 1 if (normalize_q and inv_norm_q is None) else 0
        check(lib.mcl_concept_scan_ex(q.data_ptr(), table.data_ptr(), code, Q, V, D, q.stride(0),
                                      table.stride(0), _ptr(inv_norm_q), _ptr(inv_norm_t), float(scale),
                                      float(softcap), k, index_base, _ptr(labels), val.data_ptr(),
                                      idx.data_ptr(), stats.data_ptr(), ws.data_ptr(), ws_bytes, flags,
                                      _stream(dev)))
    return val, idx, stats


# the custom op is what torch.compile / export see; eager calls go straight to the function (the
# op dispatcher costs ~15 us per call, more than half of a small-batch scan's GPU time)
_concept_scan_op = torch.library.custom_op("mcl::concept_scan", mutates_args=(), device_types="cuda")(_concept_scan_impl)


@_concept_scan_op.register_fake
def _(q, table, inv_norm_q, inv_norm_t, labels, scale, k, index_base, softcap=0.0, normalize_q=False):
    Q = q.shape[0]
    return (q.new_empty((Q, k), dtype=torch.float32), q.new_empty((Q, k), dtype=torch.int64),
            q.new_empty((Q, 4), dtype=torch.float32))


@torch.library.custom_op("mcl::merge", mutates_args=(), device_types="cuda")
def _merge_op(val: Tensor, idx: Tensor, stats: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    R, Q, k = val.shape
    dev = val.device
    val, idx, stats = val.contiguous(), idx.contiguous(), stats.contiguous()
    o_val = torch.empty((Q, k), dtype=torch.float32, device=dev)
    o_idx = torch.empty((Q, k), dtype=torch.int64, device=dev)
    o_stats = torch.empty((Q, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(load().mcl_merge(val.data_ptr(), idx.data_ptr(), stats.data_ptr(), R, Q, k,
                               o_val.data_ptr(), o_idx.data_ptr(), o_stats.data_ptr(), _stream(dev)))
    return o_val, o_idx, o_stats


@_merge_op.register_fake
def _(val, idx, stats):
    R, Q, k = val.shape
    return (val.new_empty((Q, k)), idx.new_empty((Q, k)), stats.new_empty((Q, 4)))


# --------------------------------------------------------------------------------------
# public Python API
# --------------------------------------------------------------------------------------

def row_inv_norm(x: Tensor) -> Tensor:
    """1/||row|| in fp32, zero rows -> 1 (sklearn ``normalize`` semantics)."""
    _require_cuda(x)
    _dtype_code(x)
    return torch.ops.mcl.row_inv_norm(x)


def gather_mean(table: Tensor, offsets: Tensor, ids: Tensor, normalize: bool = False,
                validate: bool = True) -> Tensor:
    """CSR gather + mean (+ L2 normalise): multi-token concept embeddings.

    ``validate`` (default): the CSR arrays are checked before the launch -- ``offsets`` must start
    at 0, be non-decreasing and end at ``len(ids)``, every id must lie in ``[0, V)`` -- and a bad
    id raises ``IndexError`` like the reference's ``embedding_matrix[token_ids]``.  Host-built CSR
    (the shims' ``tokens_to_csr``) is checked on the host for free; device-resident arrays cost one
    small reduction and a sync, which ``validate=False`` skips (ids are then clamped on the device
    and the library's bad-id flag is left unread)."""
    dev = _require_cuda(table)
    _dtype_code(table)
    if offsets.numel() < 1:
        raise ValueError("offsets must have Q+1 entries")
    if validate:
        o, i = offsets.to(torch.int64), ids.to(torch.int64)
        if int(o[0]) != 0 or int(o[-1]) != i.numel() or (o.numel() > 1 and bool((o[1:] < o[:-1]).any())):
            raise ValueError("offsets must start at 0, be non-decreasing and end at len(ids)")
        if i.numel() and (int(i.min()) < 0 or int(i.max()) >= table.shape[0]):
            raise IndexError(f"token id out of range for a table of {table.shape[0]} rows")
    offsets = offsets.to(device=dev, dtype=torch.int64).contiguous()
    ids = ids.to(device=dev, dtype=torch.int64).contiguous()
    return torch.ops.mcl.gather_mean(table, offsets, ids, bool(normalize))


@dataclass
class ScanOutput:
    topk_val: Tensor            # [Q,k] fp32, descending
    topk_idx: Tensor            # [Q,k] int64 global table rows (ties: lowest row first)
    stats: Tensor               # [Q,4] fp32: m, s, sum_z, z_label
    vocab: int                  # V the loss normalises label smoothing by
    labels: Optional[Tensor] = None
    label_smoothing: float = 0.0
    _ce: Optional[Tuple[Tensor, Tensor]] = None

    @property
    def lse(self) -> Tensor:
        return self.stats[:, 0] + torch.log(self.stats[:, 1])

    def _cross_entropy(self) -> Tuple[Tensor, Tensor]:
        """ONE library launch (``mcl_ce_from_stats``) for the per-row losses and their mean."""
        if self.labels is None:
            raise ValueError("scan was run without labels")
        if self._ce is None:
            self._ce = torch.ops.mcl.ce_from_stats(self.stats, self.labels, float(self.label_smoothing),
                                                   int(self.vocab))
        return self._ce

    @property
    def loss_rows(self) -> Tensor:
        """Per-row CE, 0 on ignored rows: (1-e)(lse - z_y) + e(lse - sum_z/V)."""
        return self._cross_entropy()[0]

    @property
    def loss(self) -> Tensor:
        """Mean over rows whose label is not -100 (``F.cross_entropy`` 'mean'; NaN if none)."""
        return self._cross_entropy()[1][0]


def concept_scan(q: Tensor, table: Tensor, k: int, *, normalize_q: bool = True,
                 normalize_t: bool = True, scale: float = 1.0, labels: Optional[Tensor] = None,
                 label_smoothing: float = 0.0, inv_norm_q: Optional[Tensor] = None,
                 inv_norm_t: Optional[Tensor] = None, index_base: int = 0,
                 vocab_total: Optional[int] = None, softcap: Optional[float] = None) -> ScanOutput:
    """Fused similarity scan of ``q [Q,D]`` against ``table [V,D]``: row-wise top-k and the
    log-sum-exp / cross-entropy statistics, without materialising the [Q,V] scores.
    ``normalize_*`` select cosine (True) vs raw dot product (False); a cached
    ``inv_norm_t`` (from :func:`row_inv_norm`) avoids re-reading the table.  ``softcap=c`` applies
    Gemma-2 style ``c * tanh(z / c)`` to every logit (HF ``final_logit_softcapping``)."""
    dev = _require_cuda(q, table, inv_norm_q, inv_norm_t)      # labels may arrive on the host
    if q.dtype != table.dtype:
        raise TypeError(f"q ({q.dtype}) and table ({table.dtype}) must have the same dtype")
    _dtype_code(q)
    if q.dim() != 2 or table.dim() != 2 or q.shape[1] != table.shape[1]:
        raise ValueError(f"shape mismatch: q {tuple(q.shape)} table {tuple(table.shape)}")
    if not 1 <= k <= min(MCL_MAX_K, table.shape[0]):
        raise ValueError(f"k={k} must be in [1, min(V, {MCL_MAX_K})]")
    if not scale > 0:
        raise ValueError("scale must be > 0")
    q, table = _rowmajor(q), _rowmajor(table)
    if normalize_t and inv_norm_t is None:
        inv_norm_t = torch.ops.mcl.row_inv_norm(table)
    if inv_norm_q is not None:
        inv_norm_q = inv_norm_q.to(torch.float32).contiguous()
    if inv_norm_t is not None:
        inv_norm_t = inv_norm_t.to(torch.float32).contiguous()
        if inv_norm_t.data_ptr() % 16:
            inv_norm_t = inv_norm_t.clone()
    if labels is not None:
        if labels.shape != (q.shape[0],):
            raise ValueError(f"labels must have shape [{q.shape[0]}]")
        if not labels.is_cuda and labels.numel() and (vocab_total is not None or index_base == 0):
            # host labels are range-checked for free; device labels are not (no sync): a label
            # outside [0, V) other than -100 then contributes z_label = 0 to its row's loss
            bad = (labels != IGNORE_INDEX) & ((labels < 0) | (labels >= int(vocab_total or table.shape[0])))
            if bool(bad.any()):
                raise IndexError("label out of range (labels are global table rows, -100 = ignore)")
        labels = labels.to(device=dev, dtype=torch.int64).contiguous()
    if softcap is not None and not softcap > 0:
        raise ValueError("softcap must be > 0 (or None)")
    scan = torch.ops.mcl.concept_scan if torch.compiler.is_compiling() else _concept_scan_impl
    val, idx, stats = scan(q, table, inv_norm_q, inv_norm_t, labels, float(scale), int(k), int(index_base),
                           float(softcap or 0.0), bool(normalize_q))
    return ScanOutput(val, idx, stats, int(vocab_total or table.shape[0]), labels,
                      float(label_smoothing))


def concept_scan_debug(q: Tensor, table: Tensor, k: int, *, inv_norm_q=None, inv_norm_t=None,
                       scale: float = 1.0, labels=None, index_base: int = 0,
                       label_smoothing: float = 0.0, softcap: float = 0.0):
    """Test hook: the same kernels, additionally dumping the score matrix [Q,V]."""
    lib = load()
    dev = _require_cuda(q, table)
    q, table = _rowmajor(q), _rowmajor(table)
    Q, D = q.shape
    V = table.shape[0]
    code = _dtype_code(q)
    val = torch.empty((Q, k), dtype=torch.float32, device=dev)
    idx = torch.empty((Q, k), dtype=torch.int64, device=dev)
    stats = torch.empty((Q, 4), dtype=torch.float32, device=dev)
    scores = torch.full((Q, V), float("nan"), dtype=torch.float32, device=dev)
    if labels is not None:
        labels = labels.to(device=dev, dtype=torch.int64).contiguous()
    with torch.cuda.device(dev):
        ws_bytes = lib.mcl_scan_workspace_bytes(Q, V, D, k, code)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        check(lib.mcl_concept_scan_softcap(q.data_ptr(), table.data_ptr(), code, Q, V, D,
                                           q.stride(0), table.stride(0), _ptr(inv_norm_q),
                                           _ptr(inv_norm_t), float(scale), float(softcap), k,
                                           index_base, _ptr(labels), val.data_ptr(), idx.data_ptr(),
                                           stats.data_ptr(), ws.data_ptr(), ws_bytes,
                                           scores.data_ptr(), _stream(dev)))
    return ScanOutput(val, idx, stats, V, labels, float(label_smoothing)), scores


def concept_scan_cta_times(q: Tensor, table: Tensor, k: int, *, inv_norm_q=None, inv_norm_t=None,
                           scale: float = 1.0):
    """Profiling hook: run the tcgen05 scan once with per-CTA globaltimer stamps and return a
    [grid, 2] int64 CPU tensor of (start, end) nanoseconds."""
    import ctypes as C
    lib = load()
    dev = _require_cuda(q, table)
    q, table = _rowmajor(q), _rowmajor(table)
    Q, D = q.shape
    V = table.shape[0]
    code = _dtype_code(q)
    with torch.cuda.device(dev):
        sm = device_info()[0]
        plan = _lib.plan_scan(Q, V, D, sm)
        grid = plan["grid"]
        val = torch.empty((Q, k), dtype=torch.float32, device=dev)
        idx = torch.empty((Q, k), dtype=torch.int64, device=dev)
        stats = torch.empty((Q, 4), dtype=torch.float32, device=dev)
        ws_bytes = lib.mcl_scan_workspace_bytes(Q, V, D, k, code)
        ws = torch.zeros(ws_bytes, dtype=torch.uint8, device=dev)
        old = lib.mcl_set_option(3, 1)
        try:
            check(lib.mcl_concept_scan(q.data_ptr(), table.data_ptr(), code, Q, V, D, q.stride(0),
                                       table.stride(0), _ptr(inv_norm_q), _ptr(inv_norm_t), float(scale),
                                       k, 0, None, val.data_ptr(), idx.data_ptr(), stats.data_ptr(),
                                       ws.data_ptr(), ws_bytes, _stream(dev)))
        finally:
            lib.mcl_set_option(3, old)
        torch.cuda.synchronize(dev)
    t = ws[: grid * 16].view(torch.int64).reshape(grid, 2).cpu()
    return t, plan


@torch.library.custom_op("mcl::similarity_matrix", mutates_args=(), device_types="cuda")
def _similarity_matrix_op(q: Tensor, table: Tensor, inv_norm_q: Optional[Tensor],
                          inv_norm_t: Optional[Tensor], scale: float) -> Tensor:
    lib = load()
    dev = q.device
    Q, D = q.shape
    V = table.shape[0]
    code = _dtype_code(q)
    out = torch.empty((Q, V), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ws_bytes = lib.mcl_similarity_workspace_bytes(Q, V, D, code)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        check(lib.mcl_similarity_matrix(q.data_ptr(), table.data_ptr(), code, Q, V, D, q.stride(0),
                                        table.stride(0), _ptr(inv_norm_q), _ptr(inv_norm_t),
                                        float(scale), out.data_ptr(), ws.data_ptr(), ws_bytes,
                                        _stream(dev)))
    return out


@_similarity_matrix_op.register_fake
def _(q, table, inv_norm_q, inv_norm_t, scale):
    return q.new_empty((q.shape[0], table.shape[0]), dtype=torch.float32)


def similarity_matrix(q: Tensor, table: Tensor, *, normalize: bool = True, scale: float = 1.0) -> Tensor:
    """Dense [Q,V] fp32 similarity matrix for SMALL problems (the all-pairs cosine matrix of
    the token analysis).  Large scans should use :func:`concept_scan`, which never stores it."""
    _require_cuda(q, table)
    if q.dtype != table.dtype:
        raise TypeError("q and table must have the same dtype")
    _dtype_code(q)
    q, table = _rowmajor(q), _rowmajor(table)
    inv_q = torch.ops.mcl.row_inv_norm(q) if normalize else None
    inv_t = torch.ops.mcl.row_inv_norm(table) if normalize else None
    return torch.ops.mcl.similarity_matrix(q, table, inv_q, inv_t, float(scale))


def merge(val: Tensor, idx: Tensor, stats: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """Merge [R,Q,k] / [R,Q,k] / [R,Q,4] per-shard results."""
    _require_cuda(val, idx, stats)
    return torch.ops.mcl.merge(val.float(), idx.to(torch.int64), stats.float())


def set_option(opt: int, value: int) -> int:
    _WS_BYTES.clear()          # plan knobs change the workspace a scan needs
    return int(load().mcl_set_option(opt, value))


def launch_count() -> int:
    return int(load().mcl_launch_count())


def device_info() -> Tuple[int, int, int]:
    import ctypes as C
    sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
    check(load().mcl_device_info(C.byref(sm), C.byref(ma), C.byref(mi)))
    return sm.value, ma.value, mi.value
