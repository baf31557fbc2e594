"""ctypes binding of libmcl_sm100.so (include/mcl.h).  No torch types cross this layer:
pointers are integers (``tensor.data_ptr()``), sizes are Python ints.

There is deliberately no fallback: if the shared library is missing the import of any
compute entry point raises, and every non-zero return code becomes an exception."""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# MCL_LIB_PATH: load an experiment build instead (A/B timing of kernel variants only)
LIB_PATH = os.environ.get("MCL_LIB_PATH") or os.path.join(_HERE, "libmcl_sm100.so")

MCL_DTYPE_BF16, MCL_DTYPE_F32 = 0, 1
MCL_MAX_K = 64
ERR_NAMES = {0: "OK", -1: "BAD_ARG", -2: "UNALIGNED", -3: "UNSUPPORTED_ARCH",
             -4: "WORKSPACE_TOO_SMALL", -5: "CUDA", -6: "NCCL", -7: "UNIMPLEMENTED"}


class MclError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libmcl_sm100: MCL_ERR_{ERR_NAMES.get(code, code)} ({code}): {msg}")
        self.code = code


_i64, _i32, _f32, _ptr, _sz = C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_size_t

# name -> (restype, argtypes); mirrors include/mcl.h line by line
SIGNATURES = {
    "mcl_version": (_i32, []),
    "mcl_last_error": (C.c_char_p, []),
    "mcl_device_info": (_i32, [C.POINTER(_i32)] * 3),
    "mcl_row_inv_norm": (_i32, [_ptr, _i32, _i64, _i64, _i64, _ptr, _ptr]),
    "mcl_gather_mean": (_i32, [_ptr, _i32, _i64, _i64, _i64, _ptr, _ptr, _i64, _i32, _ptr, _i64,
                               _ptr, _ptr]),
    "mcl_scan_workspace_bytes": (_sz, [_i64, _i64, _i64, _i32, _i32]),
    "mcl_concept_scan": (_i32, [_ptr, _ptr, _i32, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr, _f32,
                                _i32, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _sz, _ptr]),
    "mcl_concept_scan_ex": (_i32, [_ptr, _ptr, _i32, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr, _f32, _f32,
                                   _i32, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _sz, _i32, _ptr]),
    "mcl_concept_scan_softcap": (_i32, [_ptr, _ptr, _i32, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr,
                                        _f32, _f32, _i32, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _sz,
                                        _ptr, _ptr]),
    "mcl_concept_scan_debug": (_i32, [_ptr, _ptr, _i32, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr,
                                      _f32, _i32, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _sz, _ptr,
                                      _ptr]),
    "mcl_similarity_workspace_bytes": (_sz, [_i64, _i64, _i64, _i32]),
    "mcl_similarity_matrix": (_i32, [_ptr, _ptr, _i32, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr, _f32,
                                     _ptr, _ptr, _sz, _ptr]),
    "mcl_merge": (_i32, [_ptr, _ptr, _ptr, _i32, _i64, _i32, _ptr, _ptr, _ptr, _ptr]),
    "mcl_ce_from_stats": (_i32, [_ptr, _ptr, _i64, _f32, _i64, _ptr, _ptr, _ptr]),
    "mcl_ce_backward_workspace_bytes": (_sz, [_i64, _i64, _i64, _i32]),
    "mcl_ce_backward": (_i32, [_ptr, _ptr, _i32, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr, _f32, _f32, _f32,
                               _i64, _ptr, _i64, _ptr, _ptr, _ptr, _sz, _ptr]),
    "mcl_ce_backward_block_rows": (_i64, [_i64, _i64, _i32]),
    "mcl_ce_backward_ex": (_i32, [_ptr, _ptr, _i32, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr, _f32, _f32, _f32,
                                  _i64, _ptr, _i64, _ptr, _ptr, _i32, _ptr, _sz, _ptr]),
    "mcl_gemm_bf16": (_i32, [_ptr, _i32, _i64, _ptr, _i32, _i64, _ptr, _i64, _i64, _i64, _i64, _i32, _ptr]),
    "mcl_comm_unique_id": (_i32, [_ptr]),
    "mcl_comm_init": (_i32, [_ptr, _i32, _i32, C.POINTER(_ptr)]),
    "mcl_comm_destroy": (_i32, [_ptr]),
    "mcl_comm_all_gather": (_i32, [_ptr, _ptr, _ptr, _sz, _ptr]),
    "mcl_sharded_gather_bytes": (_sz, [_i64, _i32, _i32]),
    "mcl_concept_scan_sharded": (_i32, [_ptr, _ptr, _i32, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr,
                                        _f32, _i32, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _sz, _ptr,
                                        _sz, _ptr, _i32, _i32, _ptr]),
    "mcl_concept_scan_sharded_ex": (_i32, [_ptr, _ptr, _i32, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr,
                                           _f32, _i32, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _sz, _ptr,
                                           _sz, _ptr, _i32, _i32, _i32, _ptr]),
    "mcl_sharded_p2p_block_bytes": (_sz, [_i64, _i32, _i32]),
    "mcl_concept_scan_sharded_p2p": (_i32, [_ptr, _ptr, _i32, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr,
                                            _f32, _i32, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _sz,
                                            C.POINTER(_ptr), _sz, _i32, _i32, C.c_uint32, C.c_uint32, _i32, _ptr]),
    "mcl_peer_alloc": (_i32, [_sz, C.POINTER(_ptr), _ptr]),
    "mcl_peer_free": (_i32, [_ptr]),
    "mcl_peer_open": (_i32, [_ptr, C.POINTER(_ptr)]),
    "mcl_peer_close": (_i32, [_ptr]),
    "mcl_memcpy_async": (_i32, [_ptr, _ptr, _sz, _ptr]),
    "mcl_stream_wait_value32": (_i32, [_ptr, _ptr, C.c_uint32]),
    "mcl_set_option": (_i64, [_i32, _i64]),
    "mcl_launch_count": (_i64, []),
    "mcl_plan_scan": (_i32, [_i64, _i64, _i64, _i32, C.POINTER(C.c_int32)]),
    "mcl_plan_segments": (_i64, [_i64, _i64, _i64, _i32, C.POINTER(C.c_int32), _i64]),
    "mcl_plan_row_block_slots": (_i32, [_i64, _i64, _i64, _i32, _i64, C.POINTER(C.c_int32),
                                        C.POINTER(C.c_int32)]),
}

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """dlopen the library and type every entry point; raises if it is not built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} is missing. Build it with "
                    "`python -m multimodal_concept_learning_b200.build` (needs nvcc). "
                    "There is no CPU or PyTorch fallback for the concept scan.")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)   # AttributeError if the .so does not export it
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


PLAN_INTS = 88
PLAN_KEYS = ["num_rb", "num_vt", "num_kb", "cs", "workers", "gu", "ru", "waves", "S", "nslots", "grid",
             "win", "nsync", "nctr", "full_nodes", "last_nodes"]
NODE_KEYS = ["r0", "R", "a", "w0", "nfull", "tpc", "t0", "wr", "passes"]


def plan_scan(Q: int, V: int, D: int, sm: int = 148) -> dict:
    """The tcgen05 scan's tile plan (csrc/plan.h) as a dict; host-only, needs no GPU."""
    out = (C.c_int32 * PLAN_INTS)()
    check(load().mcl_plan_scan(Q, V, D, sm, out))
    v = list(out)
    plan = dict(zip(PLAN_KEYS, v))
    pos = len(PLAN_KEYS)
    for name in ("full", "last"):
        nodes = []
        for i in range(4):
            if i < plan[f"{name}_nodes"]:
                nodes.append(dict(zip(NODE_KEYS, v[pos: pos + len(NODE_KEYS)])))
            pos += len(NODE_KEYS)
        plan[name] = nodes
    return plan


def plan_segments(Q: int, V: int, D: int, sm: int = 148) -> list:
    """[(worker, row unit, first tile, end tile, slot of the unit, drift counter or -1)]."""
    lib = load()
    n = lib.mcl_plan_segments(Q, V, D, sm, None, 0)
    if n < 0:
        check(int(n))
    buf = (C.c_int32 * (6 * max(1, n)))()
    lib.mcl_plan_segments(Q, V, D, sm, buf, n)
    return [tuple(buf[6 * i: 6 * i + 6]) for i in range(n)]


def last_error() -> str:
    return (load().mcl_last_error() or b"").decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != 0:
        raise MclError(rc, last_error())
