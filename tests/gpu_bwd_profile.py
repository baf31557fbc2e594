"""Forward + backward of the fused cross-entropy at a tensor-bound shape with every row labelled, a
few times, for ncu:  python tests/gpu_bwd_profile.py [Q V D]   (default 4096 x 65536 x 1024: one row
block of the backward, eight table chunks -> per step 1 forward scan + 8 x (grad-epilogue scan,
dL/dq GEMM, dL/dT GEMM) = 25 launches of scan_tc_kernel / gemm_tc_kernel)."""
import sys
import torch
sys.path.insert(0, ".")
import multimodal_concept_learning_b200 as mcl  # noqa: E402
from multimodal_concept_learning_b200.autograd import fused_cross_entropy  # noqa: E402
Q, V, D = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (4096, 65536, 1024)
g = torch.Generator(device="cuda").manual_seed(4321)
h = (torch.randn(Q, D, generator=g, device="cuda") * 0.3).to(torch.bfloat16).requires_grad_(True)
E = (torch.randn(V, D, generator=g, device="cuda") * 0.3).to(torch.bfloat16).requires_grad_(True)
labels = torch.randint(0, V, (Q,), generator=g, device="cuda")
n0 = 0
for i in range(4):
    if i == 3:
        torch.cuda.synchronize()
        n0 = mcl.launch_count()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
    h.grad = None
    E.grad = None
    loss, _ = fused_cross_entropy(h, E, labels)
    if i == 3:
        e1.record()
    loss.backward()
e2.record()
torch.cuda.synchronize()
fl = 2.0 * Q * V * D
print(f"Q={Q} V={V} D={D}: forward {e0.elapsed_time(e1):.3f} ms ({fl / e0.elapsed_time(e1) / 1e9:.0f} TFLOP/s), "
      f"backward {e1.elapsed_time(e2):.3f} ms ({2 * fl / e1.elapsed_time(e2) / 1e9:.0f} algorithmic TFLOP/s, "
      f"{3 * fl / e1.elapsed_time(e2) / 1e9:.0f} executed), library launches per step {mcl.launch_count() - n0}, "
      f"loss {float(loss):.4f}")
