#!/usr/bin/env python3
"""Where does a scan step go?  API-level time vs memset / scan kernel / merge kernel (option 6)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multimodal_concept_learning_b200 as mcl
for name, (Q, V, D) in {"c3/8": (8192, 19008, 3584), "c2": (4096, 49408, 768), "c1": (16, 50257, 768), "c3": (8192, 152064, 3584)}.items():
    q = torch.randn(Q, D, device="cuda").bfloat16(); t = torch.randn(V, D, device="cuda").bfloat16()
    iq, it = mcl.row_inv_norm(q), mcl.row_inv_norm(t)
    for _ in range(3): mcl.concept_scan(q, t, 50, inv_norm_q=iq, inv_norm_t=it)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): mcl.concept_scan(q, t, 50, inv_norm_q=iq, inv_norm_t=it)
    e1.record(); torch.cuda.synchronize()
    api = e0.elapsed_time(e1) / 10
    mcl.set_option(6, 1)
    ph = []
    for _ in range(5):
        mcl.concept_scan(q, t, 50, inv_norm_q=iq, inv_norm_t=it)
        ph.append([mcl.set_option(100 + i, 0) / 1e6 for i in range(3)])
    mcl.set_option(6, 0)
    ph = torch.tensor(ph).median(0).values.tolist()
    print(f"{name:5s} api {api:7.3f} ms | memset {ph[0]*1e3:6.1f} us  scan {ph[1]*1e3:8.1f} us  merge {ph[2]*1e3:6.1f} us | sum {sum(ph):7.3f} ms")
    del q, t
