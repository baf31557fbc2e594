"""B200-native concept-embedding similarity scan (drop-in for the hot path of
AskSid/multimodal_concept_learning).

Layers, top to bottom::

    shims/            the reference's function names and signatures (a1-a7 of SURVEY.md section 8)
    ops.py            torch custom ops  mcl::{concept_scan, row_inv_norm, gather_mean, merge}
    graphed.py        fixed-shape scan captured in a CUDA graph (launch-bound small query batches)
    sharded.py        vocab-row sharded scan over the GPUs of one box (NCCL all-gather + merge)
    _lib.py           ctypes binding of the C ABI (include/mcl.h)
    libmcl_sm100.so   hand-written sm_100a kernels (csrc/): TMA + tcgen05 + TMEM scan,
                      merge, row-norm, gather-mean

Importing the package never touches CUDA; the first operator call loads the shared
library and fails loudly if it has not been built (there is no CPU fallback).
"""
from ._lib import MCL_MAX_K, MclError, LIB_PATH  # noqa: F401
from .ops import (ScanOutput, concept_scan, concept_scan_debug, device_info, gather_mean,  # noqa: F401
                  similarity_matrix,
                  launch_count, merge, row_inv_norm, set_option)



def __getattr__(name):            # lazy: graphed.py is only needed by callers that ask for it
    if name == "GraphedConceptScan":
        from .graphed import GraphedConceptScan
        return GraphedConceptScan
    raise AttributeError(name)


__all__ = ["concept_scan", "GraphedConceptScan", "concept_scan_debug", "similarity_matrix", "row_inv_norm", "gather_mean", "merge",
           "ScanOutput", "MclError", "MCL_MAX_K", "set_option", "launch_count", "device_info"]
