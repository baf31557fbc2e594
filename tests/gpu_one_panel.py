"""One one-row-block scan a few times, for ncu: python tests/gpu_one_panel.py Q V D [k]"""
import sys
import torch
sys.path.insert(0, ".")
import multimodal_concept_learning_b200 as mcl  # noqa: E402
Q, V, D = (int(x) for x in sys.argv[1:4])
k = int(sys.argv[4]) if len(sys.argv) > 4 else 50
q = torch.randn(Q, D, device="cuda").bfloat16()
t = torch.randn(V, D, device="cuda").bfloat16()
it = mcl.row_inv_norm(t)
for _ in range(4):
    out = mcl.concept_scan(q, t, k, inv_norm_t=it)
torch.cuda.synchronize()
print("ok", out.topk_val.shape)
