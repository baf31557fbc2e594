#!/usr/bin/env python3
"""Small-query latency breakdown: per-CTA start/end stamps of the scan kernel next to the
event-timed phases (clear / scan / merge) -- python tests/gpu_c1_diag.py [Q V D]"""
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multimodal_concept_learning_b200 as mcl  # noqa: E402
from multimodal_concept_learning_b200.ops import concept_scan_cta_times  # noqa: E402

Q, V, D = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (16, 50257, 768)
q = torch.randn(Q, D, device="cuda").bfloat16()
t = torch.randn(V, D, device="cuda").bfloat16()
iq, it = mcl.row_inv_norm(q), mcl.row_inv_norm(t)
for _ in range(20):
    mcl.concept_scan(q, t, 50, inv_norm_q=iq, inv_norm_t=it)
torch.cuda.synchronize()
for rep in range(3):
    times, plan = concept_scan_cta_times(q, t, 50, inv_norm_q=iq, inv_norm_t=it)
    t0 = int(times[:, 0].min())
    st = sorted((int(x) - t0) / 1e3 for x in times[:, 0])
    en = sorted((int(x) - t0) / 1e3 for x in times[:, 1])
    du = sorted((int(b) - int(a)) / 1e3 for a, b in times.tolist())
    print(f"grid {plan['grid']}: CTA starts us min/med/max {st[0]:.1f}/{st[len(st)//2]:.1f}/{st[-1]:.1f}  "
          f"ends {en[0]:.1f}/{en[len(en)//2]:.1f}/{en[-1]:.1f}  busy {du[0]:.1f}/{du[len(du)//2]:.1f}/{du[-1]:.1f}")
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
e[0].record()
for _ in range(200):
    mcl.concept_scan(q, t, 50, inv_norm_q=iq, inv_norm_t=it)
e[1].record()
torch.cuda.synchronize()
print(f"steady-state step {e[0].elapsed_time(e[1]) / 200 * 1e3:.1f} us")
mcl.set_option(6, 1)
ph = []
for _ in range(7):
    mcl.concept_scan(q, t, 50, inv_norm_q=iq, inv_norm_t=it)
    ph.append([mcl.set_option(100 + i, 0) / 1e3 for i in range(3)])
mcl.set_option(6, 0)
ph = torch.tensor(ph).median(0).values.tolist()
print(f"phases (events, synchronising): clear {ph[0]:.1f} us  scan {ph[1]:.1f} us  merge {ph[2]:.1f} us")
