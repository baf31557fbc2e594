#!/usr/bin/env python3
"""Small-batch scan: one-launch panel path vs the two-kernel path (library option 18), direct calls
and CUDA-graph replays, warm and after an L2 flush -- python tests/gpu_panel_timing.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multimodal_concept_learning_b200 as mcl  # noqa: E402
from multimodal_concept_learning_b200.graphed import GraphedConceptScan  # noqa: E402

HBM = 6474.0  # GB/s, MEASURED_PEAKS.json burst copy
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, n, cold=False):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    if not cold:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3


for (Q, V, D, k) in [(16, 50257, 768, 50), (96, 50257, 768, 50), (6, 50257, 768, 50), (16, 262235, 1152, 50),
                     (64, 152064, 3584, 50), (100, 152064, 3584, 50)]:
    q = torch.randn(Q, D, device="cuda").bfloat16()
    t = torch.randn(V, D, device="cuda").bfloat16()
    it = mcl.row_inv_norm(t)
    alg = 2.0 * V * D + 2.0 * Q * D + Q * (8 * k + 16)
    for two_kernel in (1, 0):
        mcl.set_option(18, two_kernel)
        direct = timed(lambda: mcl.concept_scan(q, t, k, inv_norm_t=it), 200)
        g = GraphedConceptScan(t, k, Q, inv_norm_t=it)
        g.q.copy_(q)
        graph = timed(lambda: g.graph.replay(), 500)
        cold = timed(lambda: g.graph.replay(), 30, cold=True)
        name = "two-kernel" if two_kernel else "panel     "
        print(f"Q={Q} V={V} D={D} {name}: direct {direct:.1f} us, graph {graph:.1f} us ({alg / graph / 1e3 / HBM:.2f} of HBM), "
              f"graph cold-L2 {cold:.1f} us ({alg / cold / 1e3 / HBM:.2f})", flush=True)
    mcl.set_option(18, 0)
# per-CTA stamps of the panel kernel (library option 3): start / panels done / barrier passed / end
import ctypes  # noqa: E402
from multimodal_concept_learning_b200._lib import load  # noqa: E402
lib = load()
for (Q, V, D, k) in [(16, 50257, 768, 50), (96, 50257, 768, 50), (16, 262235, 1152, 50)]:
    q = torch.randn(Q, D, device="cuda").bfloat16()
    t = torch.randn(V, D, device="cuda").bfloat16()
    iq, it = mcl.row_inv_norm(q), mcl.row_inv_norm(t)
    val = torch.empty((Q, k), dtype=torch.float32, device="cuda")
    idx = torch.empty((Q, k), dtype=torch.int64, device="cuda")
    stats = torch.empty((Q, 4), dtype=torch.float32, device="cuda")
    wsb = lib.mcl_scan_workspace_bytes(Q, V, D, k, 0)
    ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
    grid = min(148, (V + 31) // 32)
    mcl.set_option(3, 1)
    for rep in range(4):
        rc = lib.mcl_concept_scan(q.data_ptr(), t.data_ptr(), 0, Q, V, D, D, D, iq.data_ptr(), it.data_ptr(), 1.0, k, 0, None,
                                  val.data_ptr(), idx.data_ptr(), stats.data_ptr(), ws.data_ptr(), wsb,
                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
        torch.cuda.synchronize()
    mcl.set_option(3, 0)
    ts = ws[: grid * 32].view(torch.int64).reshape(grid, 4).cpu()
    t0 = int(ts[:, 0].min())
    rel = (ts - t0).double() / 1e3
    sel = rel[:, 3] > rel[:, 1] + 0.5
    print(f"Q={Q} V={V} D={D}: CTA start max {rel[:,0].max():.1f} us; panels done min/med/max "
          f"{rel[:,1].min():.1f}/{rel[:,1].median():.1f}/{rel[:,1].max():.1f}; selectors {int(sel.sum())}: barrier passed "
          f"{rel[sel,2].min():.1f}..{rel[sel,2].max():.1f}, end {rel[sel,3].min():.1f}..{rel[sel,3].max():.1f} us", flush=True)

# degenerate input (all scores equal: every key survives the bound -> exact radix select)
q0 = torch.zeros(16, 768, device="cuda").bfloat16()
t = torch.randn(50257, 768, device="cuda").bfloat16()
for two_kernel in (1, 0):
    mcl.set_option(18, two_kernel)
    g = GraphedConceptScan(t, 50, 16)
    g.q.copy_(q0)
    print("all-zero queries, two-kernel" if two_kernel else "all-zero queries, panel", f"{timed(lambda: g.graph.replay(), 100):.1f} us")
mcl.set_option(18, 0)
print("barrier faults:", mcl.set_option(104, 0))
