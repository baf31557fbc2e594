#!/usr/bin/env python3
"""Per-CTA busy time of one scan (library option 3) next to the plan's tiles per worker:
python tests/gpu_cta_times.py c3 ["7=1,8=4"]"""
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multimodal_concept_learning_b200 as mcl  # noqa: E402
from multimodal_concept_learning_b200 import _lib  # noqa: E402
from multimodal_concept_learning_b200.ops import concept_scan_cta_times  # noqa: E402
from tests.gpu_opts import SHAPES  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    cfg = [tuple(int(x) for x in kv.split("=")) for kv in (sys.argv[2] if len(sys.argv) > 2 else "").split(",") if kv]
    for o, v in cfg:
        mcl.set_option(o, v)
    Q, V, D = SHAPES[name]
    q = torch.randn(Q, D, device="cuda").bfloat16()
    t = torch.randn(V, D, device="cuda").bfloat16()
    iq, it = mcl.row_inv_norm(q), mcl.row_inv_norm(t)
    for _ in range(3):
        mcl.concept_scan(q, t, 50, inv_norm_q=iq, inv_norm_t=it)
    torch.cuda.synchronize()
    times, plan = concept_scan_cta_times(q, t, 50, inv_norm_q=iq, inv_norm_t=it)
    segs = _lib.plan_segments(Q, V, D, mcl.device_info()[0])
    tiles, nseg = {}, {}
    for w, u, a, b, j, s in segs:
        tiles[w] = tiles.get(w, 0) + b - a
        nseg[w] = nseg.get(w, 0) + 1
    t0 = int(times[:, 0].min())
    cs = plan["cs"]
    print(f"[{name} {cfg}] plan: " + " ".join(f"{k}={v}" for k, v in plan.items()))
    print(f"kernel span {(int(times[:, 1].max()) - t0) / 1e6:.3f} ms")
    rows = []
    for w in range(plan["workers"]):
        c = times[w * cs]
        rows.append((w, tiles.get(w, 0), nseg.get(w, 0), (int(c[0]) - t0) / 1e3, (int(c[1]) - int(c[0])) / 1e3))
    # group workers with the same (tiles, segments) signature
    sig = {}
    for w, nt, ns, st, du in rows:
        sig.setdefault((nt, ns), []).append(du)
    for (nt, ns), d in sorted(sig.items()):
        d = sorted(d)
        print(f"  tiles={nt:5d} segs={ns:3d} workers={len(d):3d}  busy us: min {d[0]:9.1f} med {d[len(d) // 2]:9.1f} max {d[-1]:9.1f}"
              f"  us/tile med {d[len(d) // 2] / max(1, nt):7.2f}")


if __name__ == "__main__":
    main()
