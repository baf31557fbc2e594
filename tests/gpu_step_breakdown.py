#!/usr/bin/env python3
"""Run a few scan steps of a given shape (for `ncu --metrics gpu__time_duration.sum` launch lists).
    python tests/gpu_step_breakdown.py Q V D [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multimodal_concept_learning_b200 as mcl  # noqa: E402

Q, V, D = (int(x) for x in sys.argv[1:4])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
q = torch.randn(Q, D, device="cuda").bfloat16()
t = torch.randn(V, D, device="cuda").bfloat16()
inv_q, inv_t = mcl.row_inv_norm(q), mcl.row_inv_norm(t)
for _ in range(steps):
    out = mcl.concept_scan(q, t, 50, inv_norm_q=inv_q, inv_norm_t=inv_t)
torch.cuda.synchronize()
print("ok", float(out.topk_val[0, 0]))
