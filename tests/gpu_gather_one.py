"""One gather_mean shape for ncu: python tests/gpu_gather_one.py [D] [lo] [hi]"""
import sys
import torch
sys.path.insert(0, ".")
import multimodal_concept_learning_b200 as mcl  # noqa: E402
D = int(sys.argv[1]) if len(sys.argv) > 1 else 3584
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1, 5)
g = torch.Generator(device="cuda").manual_seed(1)
V, Q = 152064, 65536
table = torch.randn(V, D, generator=g, device="cuda").to(torch.bfloat16)
lens = torch.randint(lo, hi, (Q,), generator=g, device="cuda")
offs = torch.cat([torch.zeros(1, dtype=torch.long, device="cuda"), lens.cumsum(0)])
ids = torch.randint(0, V, (int(offs[-1]),), generator=g, device="cuda")
for _ in range(5):
    out = mcl.gather_mean(table, offs, ids, False, validate=False)
torch.cuda.synchronize()
print("ok", out.shape)
