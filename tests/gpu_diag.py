#!/usr/bin/env python3
"""One-shot GPU diagnostic (not a pytest file): runs each stage of the path in a separate
subprocess with its own timeout so that a fault or hang in one stage still leaves a report
for the others.  Usage on the GPU box:

    python tests/gpu_diag.py            # all stages
    python tests/gpu_diag.py tc_small   # one stage, in-process
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = ["info", "rowops", "merge", "simt", "tc_tile", "tc_small", "tc_ragged", "tc_ties",
          "tc_perf"]
# extra stages, run by name: sched (schedule sweep + per-CTA timing)


def _imports():
    import torch
    import multimodal_concept_learning_b200 as mcl
    from oracle import concept_scan_ref as ref
    return torch, mcl, ref


def stage_info():
    torch, mcl, _ = _imports()
    print("torch", torch.__version__, "cuda", torch.version.cuda, torch.cuda.get_device_name(0))
    print("device_info", mcl.device_info())
    free, total = torch.cuda.mem_get_info()
    print(f"mem free {free/2**30:.1f} GiB / {total/2**30:.1f} GiB")


def stage_rowops():
    torch, mcl, ref = _imports()
    g = torch.Generator().manual_seed(0)
    for dt in (torch.bfloat16, torch.float32):
        x = torch.randn(1000, 776, generator=g).to(dt)
        x[3] = 0
        got = mcl.row_inv_norm(x.cuda()).cpu()
        want = ref.row_inv_norm_ref(x, torch.float64)
        print("inv_norm", dt, "max rel err", float(((got.double() - want) / want).abs().max()), "zero row ->", float(got[3]))
        offs = torch.tensor([0, 1, 4, 4, 12])
        ids = torch.randint(0, 1000, (12,), generator=g)
        for nrm in (False, True):
            got = mcl.gather_mean(x.cuda(), offs, ids, nrm).cpu()
            want = ref.gather_mean_ref(x, offs.tolist(), ids, normalize=nrm)
            print("gather_mean", dt, "normalize", nrm, "max abs diff", float((got.float() - want.float()).abs().max()),
                  "bit-exact", bool(torch.equal(got, want)))


def stage_merge():
    torch, mcl, ref = _imports()
    g = torch.Generator().manual_seed(1)
    R, Q, k = 5, 37, 50
    val = torch.randn(R, Q, k, generator=g).sort(dim=2, descending=True).values
    idx = torch.stack([torch.stack([torch.randperm(1000, generator=g)[:k] + 1000 * r for _ in range(Q)]) for r in range(R)])
    val[1, :, :5] = val[0, :, :5]   # exact cross-shard ties
    st = torch.rand(R, Q, 4, generator=g) + 0.5
    ov, oi, os_ = mcl.merge(val.cuda(), idx.cuda(), st.cuda())
    wv, wi, m, s, sz, zl = ref.merge_ref(list(val), list(idx), list(st[:, :, 0]), list(st[:, :, 1]), list(st[:, :, 2]), list(st[:, :, 3]), k)
    print("merge val exact", bool(torch.equal(ov.cpu(), wv)), "idx exact", bool(torch.equal(oi.cpu(), wi)))
    print("merge stats max err", float((os_.cpu() - torch.stack([m, s, sz, zl], 1)).abs().max()))


def _scan_report(tag, torch, mcl, ref, q, t, k, *, normalize=True, scale=1.0, labels=None, debug=True, check_exact=False):
    from tests.util import check_topk, check_stats
    qd, td = q.cuda(), t.cuda()
    inv_q = mcl.row_inv_norm(qd) if normalize else None
    inv_t = mcl.row_inv_norm(td) if normalize else None
    r = ref.concept_scan_ref(q, t, k, normalize_q=normalize, normalize_t=normalize, scale=scale, labels=labels, keep_scores=True)
    t0 = time.time()
    if debug:
        out, scores = mcl.concept_scan_debug(qd, td, k, inv_norm_q=inv_q, inv_norm_t=inv_t, scale=scale, labels=labels)
    else:
        out = mcl.concept_scan(qd, td, k, normalize_q=normalize, normalize_t=normalize, scale=scale, labels=labels, inv_norm_q=inv_q, inv_norm_t=inv_t)
        scores = None
    torch.cuda.synchronize()
    print(f"[{tag}] Q={q.shape[0]} V={t.shape[0]} D={q.shape[1]} k={k} dtype={q.dtype} ran in {time.time()-t0:.3f}s")
    if scores is not None:
        sc = scores.cpu().double()
        nan = int(torch.isnan(sc).sum())
        err = (sc - r.scores).abs()
        err[torch.isnan(err)] = float("inf")
        print(f"[{tag}] scores: nan/unwritten={nan} max abs err={float(err[~torch.isinf(err)].max()) if (~torch.isinf(err)).any() else float('nan'):.3e} "
              f"ref absmax={float(r.scores.abs().max()):.3e}")
        bad = (err > 1e-3 * max(1.0, float(r.scores.abs().max())))
        if bad.any():
            rows = bad.any(1).nonzero().flatten()
            cols = bad.any(0).nonzero().flatten()
            print(f"[{tag}] BAD scores: {int(bad.sum())} entries; rows {rows[:8].tolist()}..{rows[-3:].tolist()} n={rows.numel()}; cols {cols[:8].tolist()}..{cols[-3:].tolist()} n={cols.numel()}")
            i, j = int(rows[0]), int(cols[0])
            print(f"[{tag}] sample got {sc[i, j:j+4].tolist()} want {r.scores[i, j:j+4].tolist()}")
    try:
        check_topk(out.topk_val, out.topk_idx, r.scores, k, rtol=1e-4, atol=1e-5, exact_ties_lowest=check_exact)
        print(f"[{tag}] topk OK")
    except AssertionError as e:
        print(f"[{tag}] topk FAIL: {e}")
        print("   got idx", out.topk_idx[0, :8].tolist(), "val", [round(v, 5) for v in out.topk_val[0, :8].tolist()])
        print("   ref idx", r.topk_idx[0, :8].tolist(), "val", [round(v, 5) for v in r.topk_val[0, :8].tolist()])
    try:
        check_stats(out.stats, r, rtol=1e-4, atol=1e-4)
        print(f"[{tag}] stats OK")
    except AssertionError as e:
        print(f"[{tag}] stats FAIL: {e}")
        print("   got", out.stats[0].tolist(), "ref", [float(r.m[0]), float(r.s[0]), float(r.sum_z[0]), float(r.z_label[0])])
    return out


def stage_simt():
    torch, mcl, ref = _imports()
    from tests.util import make_inputs
    q, t = make_inputs(200, 1000, 72, 0, dtype=torch.float32)
    labels = torch.randint(0, 1000, (200,))
    labels[5] = -100
    _scan_report("simt_f32", torch, mcl, ref, q, t, 50, labels=labels)
    q, t = make_inputs(130, 5000, 64, 1, dtype=torch.bfloat16)
    mcl.set_option(2, 1)
    _scan_report("simt_bf16", torch, mcl, ref, q, t, 10, scale=20.0)
    mcl.set_option(2, 0)


def stage_tc_tile():
    """Smallest possible tcgen05 problem: one 128x256 tile, one K slice."""
    torch, mcl, ref = _imports()
    from tests.util import make_inputs
    q, t = make_inputs(128, 256, 64, 2)
    _scan_report("tc_tile_1x1x1", torch, mcl, ref, q, t, 8, normalize=False)
    q, t = make_inputs(128, 256, 256, 3)
    _scan_report("tc_tile_k4", torch, mcl, ref, q, t, 8, normalize=False)


def stage_tc_small():
    torch, mcl, ref = _imports()
    from tests.util import make_inputs
    q, t = make_inputs(256, 1024, 128, 4)
    labels = torch.randint(0, 1024, (256,))
    _scan_report("tc_small", torch, mcl, ref, q, t, 50, labels=labels)
    q, t = make_inputs(1024, 8192, 768, 5)
    _scan_report("tc_medium", torch, mcl, ref, q, t, 50, scale=100.0, labels=torch.randint(0, 8192, (1024,)))


def stage_tc_ragged():
    torch, mcl, ref = _imports()
    from tests.util import make_inputs
    q, t = make_inputs(100, 1000, 72, 6)      # Q%128, V%256, D%64 all ragged
    labels = torch.randint(0, 1000, (100,))
    labels[::7] = -100
    _scan_report("tc_ragged", torch, mcl, ref, q, t, 50, labels=labels)
    q, t = make_inputs(16, 50257, 768, 7)     # config C1 shape
    _scan_report("tc_C1", torch, mcl, ref, q, t, 50)
    for g in (1, 2, 3):
        mcl.set_option(1, g)
        q, t = make_inputs(700, 3000, 64, 8)
        _scan_report(f"tc_g{g}", torch, mcl, ref, q, t, 50, debug=False)
    mcl.set_option(1, 0)
    mcl.set_option(0, 5)   # few CTAs -> several slots (row blocks) per CTA
    q, t = make_inputs(1500, 2000, 64, 9)
    _scan_report("tc_5ctas", torch, mcl, ref, q, t, 50, debug=False)
    mcl.set_option(0, 0)


def stage_tc_ties():
    torch, mcl, ref = _imports()
    from tests.util import make_inputs
    q, t = make_inputs(64, 4096, 64, 10)
    t[2048:] = t[:2048]                        # every row duplicated: exact ties everywhere
    _scan_report("tc_dups", torch, mcl, ref, q, t, 50, check_exact=True)
    # worst case for the lazy filter: scores ascending along the table
    D = 64
    base = torch.zeros(1, D); base[0, 0] = 1.0
    t2 = (base * torch.linspace(0.01, 1.0, 6000)[:, None]).to(torch.bfloat16)
    q2 = base.repeat(40, 1).to(torch.bfloat16)
    _scan_report("tc_ascending", torch, mcl, ref, q2, t2, 50, normalize=False)
    q3 = torch.zeros(8, D, dtype=torch.bfloat16)   # zero queries: all scores 0
    _scan_report("tc_zeroq", torch, mcl, ref, q3, t, 50, check_exact=True)


def stage_tc_perf():
    torch, mcl, ref = _imports()
    for (Q, V, D) in ((4096, 49408, 768), (8192, 152064, 3584)):
        q = torch.randn(Q, D, device="cuda").bfloat16()
        t = torch.randn(V, D, device="cuda").bfloat16()
        inv_q, inv_t = mcl.row_inv_norm(q), mcl.row_inv_norm(t)
        for _ in range(2):
            out = mcl.concept_scan(q, t, 50, inv_norm_q=inv_q, inv_norm_t=inv_t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n):
            out = mcl.concept_scan(q, t, 50, inv_norm_q=inv_q, inv_norm_t=inv_t)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        tf = 2.0 * Q * V * D / (ms * 1e-3) / 1e12
        print(f"[perf] Q={Q} V={V} D={D}: {ms:.3f} ms/scan  {Q/(ms*1e-3):.3e} q/s  {tf:.1f} TFLOP/s")
        # sanity against torch on a row subset
        sub = slice(0, 64)
        z = (torch.nn.functional.normalize(q[sub].float(), dim=1) @ torch.nn.functional.normalize(t.float(), dim=1).T)
        tv, ti = torch.topk(z, 50, dim=1)
        same = (ti.sort(1).values == out.topk_idx[sub].sort(1).values).float().mean()
        print(f"[perf] top-50 index agreement with torch fp32 on 64 rows: {float(same):.4f}; "
              f"lse max err {float((torch.logsumexp(z,1) - out.lse[sub]).abs().max()):.2e}")


def _time_scan(torch, mcl, q, t, inv_q, inv_t, n=5, scale=1.0):
    for _ in range(2):
        mcl.concept_scan(q, t, 50, inv_norm_q=inv_q, inv_norm_t=inv_t, scale=scale)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        mcl.concept_scan(q, t, 50, inv_norm_q=inv_q, inv_norm_t=inv_t, scale=scale)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def stage_sched():
    """Schedule sweep: group size g and CTA count vs time, plus per-CTA busy-time spread."""
    torch, mcl, ref = _imports()
    from multimodal_concept_learning_b200.ops import concept_scan_cta_times
    shapes = {"c3": (8192, 152064, 3584), "c2": (4096, 49408, 768), "c1": (16, 50257, 768),
              "c3/8": (8192, 19008, 3584), "c4": (65536, 128256, 4096), "c5": (32768, 1048576, 1024)}
    for name, (Q, V, D) in shapes.items():
        q = torch.randn(Q, D, device="cuda").bfloat16()
        t = torch.randn(V, D, device="cuda").bfloat16()
        inv_q, inv_t = mcl.row_inv_norm(q), mcl.row_inv_norm(t)
        ms = _time_scan(torch, mcl, q, t, inv_q, inv_t)
        times, plan = concept_scan_cta_times(q, t, 50, inv_norm_q=inv_q, inv_norm_t=inv_t)
        dur = (times[:, 1] - times[:, 0]).double() / 1e3
        span = float(times[:, 1].max() - times[:, 0].min()) / 1e3
        print(f"[sched {name}] heuristic gu={plan['gu']} waves={plan['waves']} workers={plan['workers']} S={plan['S']}: {ms:.3f} ms "
              f"({2.0*Q*V*D/ms/1e9:.0f} TF/s); CTA busy us min/mean/max = {float(dur.min()):.0f}/{float(dur.mean()):.0f}/{float(dur.max()):.0f}, span {span:.0f}")
        gs = {"c3": [8, 16, 21, 32, 64], "c3/8": [8, 16, 32, 64], "c4": [37, 49, 74, 148], "c5": [18, 37, 74],
              "c2": [4, 8, 16, 32], "c1": [1]}[name]
        for g in gs:
            if g > (Q + 127) // 128:
                continue
            mcl.set_option(1, g)
            ms = _time_scan(torch, mcl, q, t, inv_q, inv_t, n=3)
            print(f"[sched {name}] g={g:3d}: {ms:.3f} ms ({2.0*Q*V*D/ms/1e9:.0f} TF/s)")
        mcl.set_option(1, 0)
        del q, t
        torch.cuda.empty_cache()


def main():
    if len(sys.argv) > 1:
        for s in sys.argv[1:]:
            globals()["stage_" + s]()
        return
    for s in STAGES:
        print(f"===== {s} =====", flush=True)
        try:
            res = subprocess.run([sys.executable, os.path.abspath(__file__), s], timeout=240,
                                 capture_output=True, text=True, cwd=ROOT)
            print(res.stdout[-6000:])
            if res.returncode != 0:
                print(f"!! stage {s} exit code {res.returncode}\n{res.stderr[-3000:]}")
        except subprocess.TimeoutExpired as e:
            print(f"!! stage {s} TIMED OUT\n{(e.stdout or b'')[-2000:]}")
        sys.stdout.flush()


if __name__ == "__main__":
    main()
