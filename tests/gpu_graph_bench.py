import sys, torch
sys.path.insert(0, '/root/repo')
import multimodal_concept_learning_b200 as mcl
for (Q,V,D) in ((16,50257,768),(96,262235,1152),(4096,49408,768)):
    q = torch.randn(Q, D, device="cuda").bfloat16(); t = torch.randn(V, D, device="cuda").bfloat16()
    it = mcl.row_inv_norm(t)
    def timeit(fn, n=200):
        for _ in range(5): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    direct = timeit(lambda: mcl.concept_scan(q, t, 50, inv_norm_t=it))
    g = mcl.GraphedConceptScan(t, 50, Q, inv_norm_t=it)
    graphed = timeit(lambda: g(q))
    print(f"Q={Q} V={V} D={D}: direct {direct:.1f} us  graphed {graphed:.1f} us")
