#!/usr/bin/env python3
"""A/B timing of two builds of the library on the same box, alternating processes so that
thermal drift cancels:  python tests/gpu_ab.py libA.so libB.so [rounds]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = {"c3": (8192, 152064, 3584), "c3/8": (8192, 19008, 3584), "c2": (4096, 49408, 768),
          "c5s": (32768, 262144, 1024), "e768": (4096, 494080, 768), "c1": (16, 50257, 768)}

CHILD = r'''
import sys, json, torch
sys.path.insert(0, %r)
import multimodal_concept_learning_b200 as mcl
res = {}
for name, (Q, V, D) in %r.items():
    q = torch.randn(Q, D, device="cuda").bfloat16(); t = torch.randn(V, D, device="cuda").bfloat16()
    iq, it = mcl.row_inv_norm(q), mcl.row_inv_norm(t)
    for _ in range(3): mcl.concept_scan(q, t, 50, inv_norm_q=iq, inv_norm_t=it)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n): mcl.concept_scan(q, t, 50, inv_norm_q=iq, inv_norm_t=it)
    e1.record(); torch.cuda.synchronize()
    res[name] = e0.elapsed_time(e1) / n
    del q, t
print(json.dumps(res))
'''


def run(lib):
    env = dict(os.environ, MCL_LIB_PATH=os.path.abspath(lib))
    out = subprocess.run([sys.executable, "-c", CHILD % (ROOT, SHAPES)], env=env, capture_output=True, text=True)
    if out.returncode != 0:
        raise SystemExit(out.stderr[-2000:])
    return json.loads(out.stdout.strip().splitlines()[-1])


def main():
    a, b = sys.argv[1], sys.argv[2]
    rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    acc = {a: {}, b: {}}
    for r in range(rounds):
        for lib in ((a, b) if r % 2 == 0 else (b, a)):
            for k, v in run(lib).items():
                acc[lib].setdefault(k, []).append(v)
    for k in SHAPES:
        ma, mb = min(acc[a][k]), min(acc[b][k])
        print(f"{k:6s} A={ma:8.3f} ms  B={mb:8.3f} ms  B/A={mb/ma:.3f}   all A={['%.3f' % x for x in acc[a][k]]} B={['%.3f' % x for x in acc[b][k]]}")


if __name__ == "__main__":
    main()
