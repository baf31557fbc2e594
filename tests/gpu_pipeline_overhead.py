#!/usr/bin/env python3
"""Host cost of one HostQueryPipeline step: e2e loop time vs device time at a shard-sized scan
(world 1), with a cProfile of the loop."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multimodal_concept_learning_b200 as mcl  # noqa: E402
from multimodal_concept_learning_b200.pipeline import HostQueryPipeline  # noqa: E402
from multimodal_concept_learning_b200.sharded import ShardedConceptScan  # noqa: E402

Q, V, D = 8192, 19008, 3584
q = torch.randn(Q, D, device="cuda").bfloat16()
t = torch.randn(V, D, device="cuda").bfloat16()
sc = ShardedConceptScan(t, V)
qh = q.cpu().pin_memory()
for use_sc in (True, False):
    pipe = HostQueryPipeline(t, 50, scanner=sc if use_sc else None)
    for _ in pipe.run(qh for _ in range(5)):
        pass
    torch.cuda.synchronize()
    n = 100
    t0 = time.perf_counter()
    for _ in pipe.run(qh for _ in range(n)):
        pass
    torch.cuda.synchronize()
    e2e = (time.perf_counter() - t0) / n
    inv = mcl.row_inv_norm(q)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        sc.scan(q, 50, inv_norm_q=inv) if use_sc else mcl.concept_scan(q, t, 50, inv_norm_q=inv, inv_norm_t=pipe.inv_norm_t)
    e1.record()
    torch.cuda.synchronize()
    print(f"scanner={use_sc}: e2e {e2e * 1e3:.3f} ms/step, device-resident step {e0.elapsed_time(e1) / n:.3f} ms")
pr = cProfile.Profile()
pipe = HostQueryPipeline(t, 50, scanner=sc)
pr.enable()
for _ in pipe.run(qh for _ in range(100)):
    pass
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
