"""The reference's LM-head training step at its native shape (1672 positions, 24 with a label, Gemma-3
table): forward + backward through fused_cross_entropy, for an ncu launch list."""
import sys
import torch
sys.path.insert(0, ".")
from multimodal_concept_learning_b200.autograd import fused_cross_entropy  # noqa: E402
Q, V, D, lab = 1672, 262235, 1152, 24
g = torch.Generator(device="cuda").manual_seed(4321)
h = (torch.randn(Q, D, generator=g, device="cuda") * 0.3).to(torch.bfloat16).requires_grad_(True)
E = (torch.randn(V, D, generator=g, device="cuda") * 0.3).to(torch.bfloat16).requires_grad_(True)
labels = torch.full((Q,), -100, dtype=torch.long, device="cuda")
rows = torch.linspace(0, Q - 1, lab, device="cuda").long()
labels[rows] = torch.randint(0, V, (lab,), generator=g, device="cuda")
for _ in range(3):
    h.grad = None
    E.grad = None
    loss, _ = fused_cross_entropy(h, E, labels)
    loss.backward()
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
h.grad = None
E.grad = None
e0.record()
loss, _ = fused_cross_entropy(h, E, labels)
e1.record()
loss.backward()
e2.record()
torch.cuda.synchronize()
print(f"forward {e0.elapsed_time(e1):.3f} ms, backward {e1.elapsed_time(e2):.3f} ms, loss {float(loss):.4f}")
