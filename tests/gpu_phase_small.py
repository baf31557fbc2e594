import os, sys, torch
sys.path.insert(0, "/root/repo")
import multimodal_concept_learning_b200 as mcl
for (Q, V, D) in [(16, 256, 64), (16, 512, 64), (16, 50257, 768), (128, 256, 64), (256, 512, 64), (16, 37888, 768)]:
    q = torch.randn(Q, D, device="cuda").bfloat16(); t = torch.randn(V, D, device="cuda").bfloat16()
    iq, it = mcl.row_inv_norm(q), mcl.row_inv_norm(t)
    for _ in range(3): mcl.concept_scan(q, t, 8, inv_norm_q=iq, inv_norm_t=it)
    mcl.set_option(6, 1)
    ph = []
    for _ in range(7):
        mcl.concept_scan(q, t, 8, inv_norm_q=iq, inv_norm_t=it)
        ph.append([mcl.set_option(100 + i, 0) / 1e3 for i in range(3)])
    mcl.set_option(6, 0)
    ph = torch.tensor(ph).median(0).values.tolist()
    from multimodal_concept_learning_b200.ops import concept_scan_cta_times
    times, plan = concept_scan_cta_times(q, t, 8, inv_norm_q=iq, inv_norm_t=it)
    span = float(times[:, 1].max() - times[:, 0].min()) / 1e3
    print(f"Q={Q} V={V} D={D}: memset {ph[0]:.1f} us scan {ph[1]:.1f} us merge {ph[2]:.1f} us | CTA span {span:.1f} us grid {plan['grid']} g={plan['g']}")
