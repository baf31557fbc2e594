"""One scan shape a few times, for ncu: python tests/gpu_one_scan.py c3/8 [k]"""
import sys
import torch
sys.path.insert(0, ".")
import multimodal_concept_learning_b200 as mcl  # noqa: E402
from tests.gpu_opts import SHAPES  # noqa: E402
Q, V, D = SHAPES[sys.argv[1] if len(sys.argv) > 1 else "c3/8"]
k = int(sys.argv[2]) if len(sys.argv) > 2 else 50
q = torch.randn(Q, D, device="cuda").bfloat16()
t = torch.randn(V, D, device="cuda").bfloat16()
iq, it = mcl.row_inv_norm(q), mcl.row_inv_norm(t)
for _ in range(4):
    out = mcl.concept_scan(q, t, k, inv_norm_q=iq, inv_norm_t=it)
torch.cuda.synchronize()
print("ok", out.topk_val.shape)
