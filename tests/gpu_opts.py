#!/usr/bin/env python3
"""A/B timing of library options on one box, interleaved in one process so that thermal and
power drift cancel:  python tests/gpu_opts.py "7=0" "7=1" "7=1,8=3" [--shapes c3,c2] [--iters 20]
Each configuration is a comma-separated list of opt=value for mcl_set_option."""
import argparse
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multimodal_concept_learning_b200 as mcl  # noqa: E402
from multimodal_concept_learning_b200 import _lib  # noqa: E402

SHAPES = {"c1": (16, 50257, 768), "c2": (4096, 49408, 768), "c3": (8192, 152064, 3584),
          "c3/2": (8192, 76032, 3584), "c3/8": (8192, 19008, 3584), "c4": (65536, 128256, 4096),
          "c5": (32768, 1048576, 1024), "c5s": (32768, 262144, 1024), "c5/8": (32768, 131072, 1024), "c4/8": (65536, 16032, 4096),
          # epilogue-bound probes: long scans (start-up transient amortised), small / CLIP-size D
          "e64": (4096, 400000, 64), "e768": (4096, 494080, 768), "e256": (4096, 400000, 256)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+")
    ap.add_argument("--shapes", default="c3,c3/8,c2,c5s")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--rounds", type=int, default=3)
    args = ap.parse_args()
    cfgs = [[tuple(int(x) for x in kv.split("=")) for kv in c.split(",") if kv] for c in args.configs]
    for name in args.shapes.split(","):
        Q, V, D = SHAPES[name]
        q = torch.randn(Q, D, device="cuda").bfloat16()
        t = torch.randn(V, D, device="cuda").bfloat16()
        iq, it = mcl.row_inv_norm(q), mcl.row_inv_norm(t)
        times = [[] for _ in cfgs]
        plans = [None] * len(cfgs)
        ref = None
        for r in range(args.rounds):
            for ci, cfg in enumerate(cfgs):
                old = [(o, mcl.set_option(o, v)) for o, v in cfg]
                try:
                    plans[ci] = _lib.plan_scan(Q, V, D, mcl.device_info()[0])
                    for _ in range(3):
                        out = mcl.concept_scan(q, t, 50, inv_norm_q=iq, inv_norm_t=it)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(args.iters):
                        out = mcl.concept_scan(q, t, 50, inv_norm_q=iq, inv_norm_t=it)
                    e1.record()
                    torch.cuda.synchronize()
                    times[ci].append(e0.elapsed_time(e1) / args.iters)
                    if ref is None:
                        ref = (out.topk_idx.clone(), out.lse.clone())
                    else:   # every configuration must give the same answer
                        same = float((out.topk_idx == ref[0]).float().mean())
                        assert same > 0.9999, f"{name} {cfg}: index agreement {same}"
                        torch.testing.assert_close(out.lse, ref[1], rtol=1e-5, atol=1e-5)
                finally:
                    for o, v in reversed(old):
                        mcl.set_option(o, v)
        flops = 2.0 * Q * V * D
        for ci, c in enumerate(args.configs):
            p = plans[ci]
            best = min(times[ci])
            print(json.dumps({"shape": name, "cfg": c, "ms_best": round(best, 4),
                              "ms_all": [round(x, 4) for x in times[ci]],
                              "tflops_best": round(flops / best / 1e9, 1),
                              "plan": {k: p[k] for k in ("cs", "workers", "gu", "waves", "S", "win", "full", "last")}}),
                  flush=True)
        del q, t
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
