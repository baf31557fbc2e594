#!/usr/bin/env python3
"""bench.py -- concept queries/sec of the fused similarity scan (top-k=50 + LSE/CE).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3]

One "step" = one pass of the hot path over one batch of synthetic queries: every query row
scored against every table row with top-k and log-sum-exp statistics produced
(`mcl_concept_scan[_sharded]`, the query row norms included).  The headline workload is
BASELINE.json configs[2] (Qwen2-VL-7B-scale vocabulary, the shape north_star quotes its target
on); with N > 1 the table is sharded by vocabulary rows across the ranks (strong scaling: total
work fixed), each step ending in one NCCL exchange + merge.  Prints ONE JSON line on rank 0.

After the timed loops (untimed) every run checks its own answer -- `parity_check`: planted exact
matches must be top-1, the sharded answer must equal an unsharded scan of the same rows (the
shards are gathered once on rank 0), and a row subsample is compared with torch fp32 -- and
exits non-zero on a mismatch.
"""
from __future__ import annotations

import argparse
import csv
import glob
import json
import os
import re
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: Q, V, D, cosine?, scale, labels?, top-k, BASELINE.json config index
    "c1": dict(Q=16, V=50257, D=768, normalize=True, scale=1.0, labels=False, k=50, cfg=0,
               desc="16 concept embeddings x GPT-2 vocab 50257x768, cosine top-50"),
    "c2": dict(Q=4096, V=49408, D=768, normalize=True, scale=100.0, labels=True, k=50, cfg=1,
               desc="4096 queries x CLIP ViT-L/14 text table 49408x768, top-50 + softmax-CE"),
    "c3": dict(Q=8192, V=152064, D=3584, normalize=True, scale=1.0, labels=False, k=50, cfg=2,
               desc="8192 multi-token concept embeddings x Qwen2-VL-7B vocab 152064x3584, cosine top-50 + LSE"),
    "c4": dict(Q=65536, V=128256, D=4096, normalize=True, scale=1.0, labels=False, k=50, cfg=3,
               desc="65536 queries x Llama-3-8B vocab 128256x4096, top-50 + LSE"),
    "c5": dict(Q=32768, V=1048576, D=1024, normalize=True, scale=100.0, labels=True, k=50, cfg=4,
               desc="32768 image embeddings x 1M concept bank 1048576x1024, contrastive logits + CE"),
    # the reference's own LM head (src/multimodal/mllm.py:115 at batch 8: 8 x 209 positions against the
    # Gemma-3 table after the OOD tokens were added), raw dot product, CE + argmax (multimodal_training.py:276)
    "gemma3_head": dict(Q=1672, V=262235, D=1152, normalize=False, scale=1.0, labels=True, k=1, cfg=None,
                        desc="reference-native LM head: 1672 hidden states x Gemma-3 table 262235x1152, CE + argmax (k=1)"),
    # ... and its evaluation with label-aware row selection (shims/mllm.py rows="labelled"): the 3 answer
    # positions of each of 8 samples (multimodal_training.py:276-285 reads no other row), one row block
    "gemma3_eval": dict(Q=24, V=262235, D=1152, normalize=False, scale=1.0, labels=True, k=1, cfg=None,
                        desc="reference-native evaluation, labelled rows only: 24 hidden states x Gemma-3 table 262235x1152, CE + argmax (k=1)"),
}
K_TOP = 50
METRIC = "concept queries/sec vs vocab (top-k=50)"
NEAR_TIE_GAP = 1e-3


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update({k: float(m[k]) for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained") if k in m})
        p["source"] = "measured"
    except Exception:
        pass
    return p


_UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def ncu_traffic(workload, world, kernel="scan_tc"):
    """`roofline.traffic`: dram__bytes_read.sum + dram__bytes_write.sum of the scan kernel, per
    launch, from the newest committed `ncu --set full` raw page of this workload
    (profiles/r*_<workload>[_nN]_<kernel>*_ncu_raw.csv).  None when no capture exists."""
    tag = workload if world == 1 else f"{workload}_n{world}"
    files = [f for f in glob.glob(os.path.join(ROOT, "profiles", f"r*_{tag}_{kernel}*_ncu_raw.csv"))
             if world > 1 or not re.search(r"_n\d+_", os.path.basename(f))]
    if not files:
        return None, None
    path = sorted(files, key=lambda f: (os.path.basename(f).split("_")[0], os.path.getmtime(f)))[-1]
    try:
        rows = list(csv.reader(open(path)))
        head, units = rows[0], rows[1]
        ik, ir, iw = head.index("Kernel Name"), head.index("dram__bytes_read.sum"), head.index("dram__bytes_write.sum")
        # (the main scan kernel: epilogue modes 0 / 1; mode 2 is the threshold-seeding pre-pass)
        vals = [float(r[ir]) * _UNIT[units[ir]] + float(r[iw]) * _UNIT[units[iw]] for r in rows[2:]
                if re.search(r"scan_tc_kernel<\d, \d(, [01])?>|panel_scan_kernel", r[ik])]
        if not vals:
            return None, None
        return sum(vals) / len(vals), os.path.relpath(path, ROOT)
    except Exception:
        return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (rank 0)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


def make_inputs(w, device, rank, world, seed_base=1234):
    """Synthetic random-init embeddings of the named shape; each rank builds only its shard."""
    from multimodal_concept_learning_b200.sharded import shard_rows
    cfg = w["cfg"] if w["cfg"] is not None else 9
    g = torch.Generator(device=device).manual_seed(seed_base + cfg)
    q = torch.randn(w["Q"], w["D"], generator=g, device=device).to(torch.bfloat16)
    labels = torch.randint(0, w["V"], (w["Q"],), generator=g, device=device) if w["labels"] else None
    if w["k"] == 1 and labels is not None and w["Q"] >= 209:
        # answer-only supervision (imagenet_dataset.py:171-175): 1-3 labelled positions per sample of 209
        keep = torch.zeros(w["Q"], dtype=torch.bool, device=device)
        keep[torch.arange(205, w["Q"], 209, device=device)] = True
        keep[torch.arange(206, w["Q"], 209, device=device)] = True
        labels = torch.where(keep, labels, torch.full_like(labels, -100))
    lo, hi = shard_rows(w["V"], world, rank)
    gt = torch.Generator(device=device).manual_seed(seed_base + 100 * (rank + 1) + cfg)
    table = torch.randn(hi - lo, w["D"], generator=gt, device=device).to(torch.bfloat16)
    return q, table, labels, lo, hi


def time_steps(fn, steps, warmup, world, device):
    """W untimed steps, then exactly K steps between barrier+synchronize, CUDA events on the
    launching (current) stream; returns max-over-ranks milliseconds for the K steps."""
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


def build_step(w, q, table, labels, lo, world, vocab_total):
    """Returns (step_fn, close_fn): the public-API call a user makes, inputs resident.  The table's
    inverse row norms are cached per table version (4 B/row); everything that depends on the query
    batch -- its row norms included -- runs inside the step."""
    import multimodal_concept_learning_b200 as mcl
    if world > 1:
        from multimodal_concept_learning_b200.sharded import ShardedConceptScan
        # result exchange: peer-memory stores over NVLink by default ("auto"); MCL_SHARDED_EXCHANGE=nccl
        # measures the NCCL path (grouped send/recv + all-gather) for comparison
        sc = ShardedConceptScan(table, vocab_total, normalize_t=w["normalize"],
                                exchange=os.environ.get("MCL_SHARDED_EXCHANGE", "auto"))

        def step(qq=q):
            return sc.scan(qq, w["k"], normalize_q=w["normalize"], scale=w["scale"], labels=labels)
        step.scanner, step.inv_t = sc, None
        return step, sc.close
    inv_t = mcl.row_inv_norm(table) if w["normalize"] else None   # cached per table version

    def step(qq=q):
        return mcl.concept_scan(qq, table, w["k"], normalize_q=w["normalize"], normalize_t=w["normalize"],
                                scale=w["scale"], labels=labels, inv_norm_t=inv_t)
    step.scanner, step.inv_t = None, inv_t
    return step, (lambda: None)


def topk_matches_reference(val, idx, z, k, rtol=1e-4, gap=NEAR_TIE_GAP):
    """ours (val, idx) [R,k] against reference scores z [R,V] (torch fp32, same device): values at
    our indices agree, nothing we return is below the near-tie band of the true k-th value, and
    every reference item clearly above the boundary is returned."""
    got = torch.gather(z, 1, idx)
    scale = z.abs().max().clamp_min(1.0)
    val_err = float(((val - got).abs() / (got.abs() + 1e-3 * scale)).max())
    rv, ri = torch.topk(z, min(k + 1, z.shape[1]), dim=1)
    kth = rv[:, k - 1:k]
    nxt = rv[:, k:k + 1] if rv.shape[1] > k else torch.full_like(kth, -float("inf"))
    g = gap * float(scale)
    inside = bool((got >= kth - g).all())
    clear = rv[:, :k] > nxt + g
    present = (ri[:, :k, None] == idx[:, None, :]).any(-1)
    complete = bool((present | ~clear).all())
    return val_err, bool(inside and complete and val_err <= 10 * rtol)


def parity_check(w, step, q, table, labels, lo, hi, world, rank, device, rows=256, planted=64):
    """Untimed self-check of the answer the timed loops computed (see module docstring)."""
    import torch.distributed as dist
    import multimodal_concept_learning_b200 as mcl
    V, D, k = w["V"], w["D"], w["k"]
    rows = min(rows, w["Q"])
    planted = min(planted, rows)
    # planted exact matches: query row i becomes a copy of global table row ids[i] (x8 for raw dot
    # products so that <t, 8t> dominates); the rank that owns the row provides it
    ids = torch.linspace(0, V - 1, planted, device=device).long()
    rowsbuf = torch.zeros(planted, D, dtype=torch.float32, device=device)
    mine = (ids >= lo) & (ids < hi)
    rowsbuf[mine] = table[ids[mine] - lo].float()
    if world > 1:
        dist.all_reduce(rowsbuf)
    qs = q[:rows].clone()
    qs[:planted] = (rowsbuf * (1.0 if w["normalize"] else 8.0)).to(torch.bfloat16)
    lab = labels[:rows].clone() if labels is not None else None
    if lab is not None:
        lab[:planted] = ids
    out = (step.scanner.scan(qs, k, normalize_q=w["normalize"], scale=w["scale"], labels=lab) if world > 1
           else mcl.concept_scan(qs, table, k, normalize_q=w["normalize"], normalize_t=w["normalize"],
                                 scale=w["scale"], labels=lab, inv_norm_t=step.inv_t))
    res = {"rows": rows, "planted": planted}
    # the whole table on rank 0
    full = table
    if world > 1:
        from multimodal_concept_learning_b200.sharded import shard_rows
        if rank == 0:
            full = torch.empty(V, D, dtype=table.dtype, device=device)
            full[lo:hi] = table
            for r in range(1, world):
                rlo, rhi = shard_rows(V, world, r)
                dist.recv(full[rlo:rhi], src=r)
        else:
            dist.send(table.contiguous(), dst=0)
    if rank == 0:
        res["planted_top1"] = bool(torch.equal(out.topk_idx[:planted, 0], ids))
        if w["normalize"]:
            res["planted_top1"] = res["planted_top1"] and bool(
                ((out.topk_val[:planted, 0] - w["scale"]).abs() <= 1e-4 * w["scale"]).all())
        if world > 1:
            un = mcl.concept_scan(qs, full, k, normalize_q=w["normalize"], normalize_t=w["normalize"],
                                  scale=w["scale"], labels=lab)
            differ = out.topk_idx != un.topk_idx
            # a tie AT the k-th value between distinct rows may be resolved differently by the two
            # partitions of the columns; anything else must agree bit for bit
            res["idx_equal"] = bool(not (differ & (un.topk_val != un.topk_val[:, -1:])).any()
                                    and float(differ.float().mean()) < 1e-3)
            res["val_equal"] = bool(torch.equal(out.topk_val, un.topk_val))
            res["lse_max_rel"] = float(((out.lse - un.lse).abs() / un.lse.abs().clamp_min(1e-6)).max())
            if lab is not None:
                res["loss_rel"] = float((out.loss - un.loss).abs() / un.loss.abs().clamp_min(1e-6))
        fn = torch.nn.functional.normalize
        z = (fn(qs.float(), dim=1) @ fn(full.float(), dim=1).T if w["normalize"] else qs.float() @ full.float().T) * w["scale"]
        res["val_max_rel_vs_torch_fp32"], res["topk_ok_vs_torch_fp32"] = topk_matches_reference(
            out.topk_val, out.topk_idx, z, k)
        lse_ref = torch.logsumexp(z, 1)
        res["lse_max_rel_vs_torch_fp32"] = float(((out.lse - lse_ref).abs() / lse_ref.abs().clamp_min(1e-6)).max())
        if lab is not None:
            want = torch.nn.functional.cross_entropy(z, lab)
            res["loss_rel_vs_torch_fp32"] = float((out.loss - want).abs() / want.abs().clamp_min(1e-6))
        res["ok"] = bool(res["planted_top1"] and res["topk_ok_vs_torch_fp32"]
                         and res["lse_max_rel_vs_torch_fp32"] <= 1e-4
                         and res.get("loss_rel_vs_torch_fp32", 0.0) <= 1e-4
                         and res.get("idx_equal", True) and res.get("val_equal", True)
                         and res.get("lse_max_rel", 0.0) <= 1e-5)
        del z
    del full
    return res


def gpu_unfused_baseline(w, q, table, labels, steps, device):
    """The reference's own composition run on THIS GPU (what src/multimodal/mllm.py:115 computes on a
    GPU, SURVEY section 8d "honest GPU comparator"): cuBLAS bf16 `@` -> fp32 upcast -> torch.topk ->
    logsumexp / F.cross_entropy, chunked over the queries so that the [chunk, V] logits fit.
    The normalised table is prepared once outside the timed region, as our arm caches 1/||row||."""
    fn = torch.nn.functional.normalize
    tn = fn(table.float(), dim=1).to(torch.bfloat16) if w["normalize"] else table
    chunk = max(128, min(w["Q"], int(3.0e9 // (6 * w["V"]))))

    def step():
        outs = []
        for a in range(0, w["Q"], chunk):
            qc = q[a:a + chunk]
            if w["normalize"]:
                qc = fn(qc.float(), dim=1).to(torch.bfloat16)
            z = (qc @ tn.T).float() * w["scale"]                 # bf16 GEMM, fp32 upcast (loss_utils.py:55)
            val, idx = torch.topk(z, w["k"], dim=1)
            if labels is not None:
                outs.append((val, idx, torch.nn.functional.cross_entropy(z, labels[a:a + chunk], reduction="sum")))
            else:
                outs.append((val, idx, torch.logsumexp(z, dim=1)))
        return outs
    for _ in range(2):
        step()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / steps
    del tn
    return {"value": w["Q"] / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms, "steps": steps,
            "what": "torch on the same GPU: cuBLAS bf16 matmul -> fp32 logits -> topk -> logsumexp/cross_entropy, "
                    f"chunks of {chunk} queries; normalised table cached outside the timed region"}


def row_kernel_rooflines(w, table, device, pk, steps=20):
    """Part (c) of the path (HBM-bound row kernels) on the headline table: algorithmic bytes /
    CUDA-event time against the measured copy bandwidth."""
    import multimodal_concept_learning_b200 as mcl
    out = []

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(device)
        return e0.elapsed_time(e1) / steps
    V, D = table.shape
    ms = timed(lambda: mcl.row_inv_norm(table))
    b = V * D * 2.0 + V * 4.0
    out.append({"kernel": "row_inv_norm_kernel", "rows": V, "dim": D, "ms": ms, "alg_bytes": b,
                "hbm_gbs": b / ms / 1e6, "hbm_frac": b / ms / 1e6 / pk["hbm_gbs"],
                "l2": "table larger than L2" if V * D * 2 > 126e6 else "fits in L2"})
    # multi-token concept embeddings (a3): Q rows, 1-4 random table rows each, mean + L2 normalise
    g = torch.Generator(device=device).manual_seed(77)
    Q = 65536
    lens = torch.randint(1, 5, (Q,), generator=g, device=device)
    offs = torch.cat([torch.zeros(1, dtype=torch.long, device=device), lens.cumsum(0)])
    ids = torch.randint(0, V, (int(offs[-1]),), generator=g, device=device)
    ms = timed(lambda: mcl.gather_mean(table, offs, ids, True, validate=False))
    b = ids.numel() * D * 2.0 + Q * D * 2.0 + ids.numel() * 8.0 + (Q + 1) * 8.0
    out.append({"kernel": "gather_mean_kernel", "rows": Q, "nnz": int(ids.numel()), "dim": D, "ms": ms,
                "alg_bytes": b, "hbm_gbs": b / ms / 1e6, "hbm_frac": b / ms / 1e6 / pk["hbm_gbs"],
                "l2": "gathered rows + output larger than L2"})
    return out


def run_workload(name, steps, warmup, world, rank, device, with_e2e=True, sampler=None, check=True):
    import torch.distributed as dist
    import multimodal_concept_learning_b200 as mcl
    w = WORKLOADS[name]
    q, table, labels, lo, hi = make_inputs(w, device, rank, world)
    step, close = build_step(w, q, table, labels, lo, world, w["V"])
    if sampler:
        sampler.start()
    n0 = mcl.launch_count()
    ms = time_steps(step, steps, warmup, world, device)
    launches = (mcl.launch_count() - n0) * steps // (steps + warmup)
    clocks = sampler.stop() if sampler else None
    res = {"ms_per_step": ms / steps, "value": w["Q"] * steps / (ms * 1e-3), "gpu_launches": int(launches),
           "clocks": clocks}
    if world == 1 and 2.0 * w["V"] * w["D"] < 126e6:
        # inputs that fit in the 126 MB L2 (C1, C2): the back-to-back loop above finds the table in
        # L2; also time every step alone after a 256 MB write has flushed it (cold-L2 number)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
        n_cold = max(5, steps)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_cold)]
        for e0, e1 in evs:
            flush.zero_()
            e0.record()
            step()
            e1.record()
        torch.cuda.synchronize(device)
        res["cold_ms_per_step"] = sum(e0.elapsed_time(e1) for e0, e1 in evs) / n_cold
        del flush
    if world == 1 and w["Q"] <= 128:
        # launch-bound regime: the same step replayed from a CUDA graph (public API GraphedConceptScan)
        g = mcl.GraphedConceptScan(table, w["k"], w["Q"], normalize=w["normalize"], scale=w["scale"],
                                   inv_norm_t=step.inv_t, with_labels=labels is not None)
        gms = time_steps(lambda: g(q, labels), steps, warmup, world, device)
        res["graphed_ms_per_step"] = gms / steps
        # ... and with the batch already in the graph's own buffer (no input copy: the bare replay)
        g.q.copy_(q)
        if labels is not None:
            g.labels.copy_(labels)
        gms0 = time_steps(lambda: g(g.q, g.labels), steps, warmup, world, device)
        res["graphed_inplace_ms_per_step"] = gms0 / steps
        del g
    flops = 2.0 * w["Q"] * w["V"] * w["D"]
    res["tflops"] = flops / (ms / steps * 1e-3) / 1e12
    res["alg_bytes"] = 2.0 * (w["V"] * w["D"] + w["Q"] * w["D"]) + w["Q"] * (8 * w["k"] + 16)
    if with_e2e:
        # same metric through the public API with HOST buffers: every step copies the query batch
        # from pinned host memory and returns (top-k values, indices, stats) in host memory.
        # HostQueryPipeline overlaps the copies of neighbouring steps with the scan.
        from multimodal_concept_learning_b200.pipeline import HostQueryPipeline
        q_host = q.cpu().pin_memory()
        # (N > 1: every rank uploads 1/N of the batch, pushes it to the peers over NVLink with copy
        # engines and takes home the 1/N of the result rows it merged)
        pipe = HostQueryPipeline(table, w["k"], normalize=w["normalize"], scale=w["scale"],
                                 inv_norm_t=step.inv_t, scanner=step.scanner, reuse_host_buffers=True,
                                 local_rows=world > 1)
        outs = None
        # (warm-up long enough for the pinned result buffers of all batches in flight to come from
        # torch's host-allocator cache: a cudaHostAlloc inside the timed region costs milliseconds)
        for outs in pipe.run((q_host for _ in range(max(8, warmup))), labels):
            pass
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for outs in pipe.run((q_host for _ in range(steps)), labels):
            pass                                           # the caller holds the result on the host
        torch.cuda.synchronize(device)
        dt = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        res["e2e"] = {"value": w["Q"] * steps / float(dt), "unit": "queries/s",
                      "ms_per_step": 1e3 * float(dt) / steps,
                      "h2d_bytes_per_step": pipe.h2d_bytes_per_step(q_host),
                      # summed over the ranks (each copies the rows it merged; all rows at N <= 2)
                      "d2h_bytes_per_step": world * sum(t.numel() * t.element_size() for t in outs),
                      "note": pipe.describe()}
        if world > 1 and pipe._board is not None:
            # where a step's time goes: the query distribution alone (upload of 1/N + pushes to the
            # peers + the wait for every peer's slice), device-timed, without any scan behind it
            board, cs = pipe._board, pipe.copy_stream
            main = torch.cuda.current_stream(device)
            for _ in range(3):
                board.wait(*board.publish(q_host, cs), main)
            torch.cuda.synchronize(device)
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main)
            for _ in range(steps):
                board.wait(*board.publish(q_host, cs), main)
            e1.record(main)
            torch.cuda.synchronize(device)
            dms = torch.tensor([e0.elapsed_time(e1) / steps], device=device, dtype=torch.float64)
            dist.all_reduce(dms, op=dist.ReduceOp.MAX)
            res["e2e"]["breakdown_ms"] = {"scan_exchange_merge": ms / steps, "distribute_queries_alone": float(dms),
                                          "e2e_step": 1e3 * float(dt) / steps}
        pipe.close()
    if check:
        res["parity_check"] = parity_check(w, step, q, table, labels, lo, hi, world, rank, device)
    res["inputs"] = (q, table, labels)
    res["close"] = close
    return res


def release(res):
    res.pop("close")()
    res.pop("inputs", None)
    torch.cuda.empty_cache()


def cpu_reference_step(w, n_queries, seed=1234):
    """The reference's own CPU path for this workload -- F.normalize -> @ -> topk -> logsumexp /
    cross_entropy in fp32 (oracle/concept_scan_ref.torch_composition_ref) -- on a bounded
    query sample against the FULL table; returns seconds for one step.  The fp32 / normalised copy
    of the table is prepared once per table, outside the timed step, exactly as the GPU arm caches
    the table's inverse norms: both arms time the work that depends on the query batch."""
    from oracle.concept_scan_ref import prepare_table_ref, torch_composition_ref
    cfg = w["cfg"] if w["cfg"] is not None else 9
    g = torch.Generator().manual_seed(seed + cfg)
    q = torch.randn(n_queries, w["D"], generator=g).to(torch.bfloat16)
    cache = cpu_reference_step.__dict__.setdefault("tables", {})
    key = (w["V"], w["D"], w["normalize"])
    if key not in cache:
        cache.clear()
        cache[key] = prepare_table_ref(torch.randn(w["V"], w["D"], generator=g).to(torch.bfloat16), w["normalize"])
    table = cache[key]
    labels = torch.randint(0, w["V"], (n_queries,), generator=g) if w["labels"] else None
    t0 = time.perf_counter()
    torch_composition_ref(q, table, w["k"], normalize=w["normalize"], scale=w["scale"], labels=labels,
                          table_prepared=True)
    return time.perf_counter() - t0


def cpu_sample_size(w):
    # a few seconds of CPU work per step on a typical host (~1 TFLOP/s fp32 over all cores); the
    # [n, V] fp32 score matrix bounds n as well (<= 8 GB)
    n = int(3.0e12 / (2.0 * w["V"] * w["D"]))
    n = min(n, int(8e9 / (4.0 * w["V"])))
    return max(16, min(w["Q"], n))


CPU_SAMPLE_NOTE = ("torch fp32 normalize(q) -> matmul -> topk -> logsumexp/CE on all host threads; "
                   "fp32 normalised table prepared once outside the timed step")


def backward_bench(Q, V, D, labelled, device, pk, steps=3):
    """forward + backward of the fused cross-entropy through the public autograd API
    (`fused_cross_entropy(...).backward()`): the scan's grad epilogue recomputes dL/dz per tile, two
    tcgen05 GEMMs form dL/dq and dL/dT.  `labelled` rows carry a label (the others are -100)."""
    from multimodal_concept_learning_b200.autograd import fused_cross_entropy
    import multimodal_concept_learning_b200 as mcl
    g = torch.Generator(device=device).manual_seed(4321)
    h = (torch.randn(Q, D, generator=g, device=device) * 0.3).to(torch.bfloat16).requires_grad_(True)
    E = (torch.randn(V, D, generator=g, device=device) * 0.3).to(torch.bfloat16).requires_grad_(True)
    labels = torch.full((Q,), -100, dtype=torch.long, device=device)
    rows = torch.linspace(0, Q - 1, labelled, device=device).long()
    labels[rows] = torch.randint(0, V, (labelled,), generator=g, device=device)

    def step():
        h.grad = None
        E.grad = None
        loss, _ = fused_cross_entropy(h, E, labels)
        loss.backward()
        return loss
    step()
    torch.cuda.synchronize(device)
    n0 = mcl.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / steps
    fwd = 2.0 * Q * V * D                       # the forward scans every row
    bwd_alg = 4.0 * labelled * V * D            # dL/dq and dL/dT
    bwd_exec = 6.0 * labelled * V * D           # + the recomputed scores
    # bytes the backward cannot avoid: the table once, the fp32 table gradient once
    hbm = (2.0 * V * D + 4.0 * V * D) / (ms * 1e-3) / 1e9
    out = {"Q": Q, "V": V, "D": D, "labelled_rows": labelled, "ms_fwd_bwd": ms,
           "alg_tflops": (fwd + bwd_alg) / (ms * 1e-3) / 1e12,
           "tensor_frac": (fwd + bwd_alg) / (ms * 1e-3) / 1e12 / pk["bf16_tflops"],
           "executed_tflops": (fwd + bwd_exec) / (ms * 1e-3) / 1e12,
           "min_hbm_gbs": hbm, "min_hbm_frac": hbm / pk["hbm_gbs"],
           "gpu_launches": int((mcl.launch_count() - n0) // steps),
           "what": "fused_cross_entropy(h, E, labels) + loss.backward(): tcgen05 forward scan (k=1), grad-epilogue "
                   "scan + two tcgen05 GEMMs per dL/dz block (few rows: dL/dq split over K, table gradient written in "
                   "bf16); includes the torch glue (row gather, casts)"}
    del h, E
    torch.cuda.empty_cache()
    return out


def literal_pair_loop_baseline(device_unused=None):
    """BASELINE.md section 4.1 / token_embedding_analysis.py:237-246: the reference's literal
    per-pair sklearn loop for the 16 x 16 self-similarity of the C1 concept embeddings, and the
    batched sklearn call + topk against the whole GPT-2-size table."""
    import numpy as np
    from oracle.reference_sites import pairwise_cosine_distance_sklearn_loop
    from sklearn.metrics.pairwise import cosine_similarity
    w = WORKLOADS["c1"]
    g = torch.Generator().manual_seed(1234)
    e = torch.randn(w["Q"], w["D"], generator=g).numpy().astype(np.float32)
    t = torch.randn(w["V"], w["D"], generator=g).numpy().astype(np.float32)
    pairwise_cosine_distance_sklearn_loop(e[:4])
    t0 = time.perf_counter()
    d = pairwise_cosine_distance_sklearn_loop(e)
    t_loop = time.perf_counter() - t0
    t0 = time.perf_counter()
    z = cosine_similarity(e, t)
    torch.topk(torch.from_numpy(z), w["k"], dim=1)
    t_batched = time.perf_counter() - t0
    return {"pairs": int(len(d)), "us_per_pair": 1e6 * t_loop / len(d), "loop_ms": 1e3 * t_loop,
            "batched_cosine_topk_ms": 1e3 * t_batched, "batched_queries_per_s": w["Q"] / t_batched,
            "what": "sklearn cosine_similarity([a],[b]) per pair i<j of the 16 concept embeddings (literal "
                    "reference loop); then ONE cosine_similarity(q, table) + torch.topk(50) against 50257x768"}


def main():
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner)
    # are sent to stderr for the duration of the run
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-sweep", action="store_true", help="skip the secondary workloads (N=1 only)")
    ap.add_argument("--profile", action="store_true",
                    help="kernel-only run for ncu: no sweep, no e2e, no CPU baseline, no check (not a bench value)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3                                   # timing rule: W >= 3
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w = WORKLOADS[args.workload]
    pk = peaks()

    cfg_name = f"BASELINE configs[{w['cfg']}]" if w["cfg"] is not None else "reference-native shape"
    base = {"metric": METRIC if w["k"] == K_TOP else f"concept queries/sec vs vocab (top-k={w['k']})",
            "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{cfg_name}: {w['desc']}", "Q": w["Q"],
                       "V": w["V"], "D": w["D"], "k": w["k"],
                       "parallelism": (f"vocab-row sharding x{args.gpus}, result exchange: "
                                       + ("NCCL" if os.environ.get("MCL_SHARDED_EXCHANGE") == "nccl" else "peer-memory stores over NVLink")
                                       ) if args.gpus > 1 else "single GPU",
                       "l2": ("inputs larger than L2 (table {:.2f} GB vs 126 MB)" if w["V"] * w["D"] * 2 > 126e6 else
                              "table {:.2f} GB fits in the 126 MB L2: ms_per_step is warm, roofline.cold_l2 after a 256 MB flush"
                              ).format(w["V"] * w["D"] * 2 / 1e9)}}

    if args.impl == "reference":
        if rank != 0:
            return
        torch.set_num_threads(os.cpu_count() or 1)
        n = cpu_sample_size(w)
        for _ in range(min(args.warmup, 1)):
            cpu_reference_step(w, n)
        steps = max(1, min(args.steps, 3))
        t = sum(cpu_reference_step(w, n) for _ in range(steps))
        v = n * steps / t
        base.update({"impl": "reference", "value": v, "ms_per_step": 1e3 * t / steps, "steps": steps,
                     "dtype": "f32", "n_gpus": args.gpus, "gpu_launches": 0,
                     "cpu_baseline": {"value": v, "unit": "queries/s", "cores": torch.get_num_threads(),
                                      "kind": "port",
                                      "sample": f"{n} of {w['Q']} queries per step against the full table; " + CPU_SAMPLE_NOTE},
                     "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        emit(base)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    import multimodal_concept_learning_b200 as mcl
    mcl.device_info()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    res = run_workload(args.workload, args.steps, args.warmup, world, rank, device, not args.profile, sampler,
                       check=not args.profile)
    out = dict(base)
    out.update({"value": res["value"], "ms_per_step": res["ms_per_step"], "e2e": res.get("e2e"),
                "gpu_launches": res["gpu_launches"], "clocks": res["clocks"],
                "parity_check": res.get("parity_check")})
    shard_flops = 2.0 * w["Q"] * w["V"] * w["D"] / world
    achieved = shard_flops / (res["ms_per_step"] * 1e-3) / 1e12
    traffic, traffic_src = ncu_traffic(args.workload, world, "panel_scan" if w["Q"] <= 128 else "scan_tc")
    alg_bytes = 2.0 * (w["V"] / world * w["D"] + w["Q"] * w["D"]) + w["Q"] * (8 * w["k"] + 16)
    if w["Q"] <= 128:
        # one row block: AI ~ Q flop/B, the table read bounds the scan (panel_scan.cu, ONE launch per step);
        # `achieved` from the direct-call step time, `graphed` from CUDA-graph replays of the same step
        gbs = alg_bytes / (res["ms_per_step"] * 1e-3) / 1e9
        out["roofline"] = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                           "frac": gbs / pk["hbm_gbs"], "peak_source": pk["source"] + " (burst copy)",
                           "traffic": traffic, "traffic_source": traffic_src,
                           "traffic_unit": "bytes per launch (algorithmic: %.3e)" % alg_bytes,
                           "kernel": "panel_scan_kernel (per GPU; the step is this one launch)"}
        if "graphed_ms_per_step" in res:
            gg = alg_bytes / (res["graphed_ms_per_step"] * 1e-3) / 1e9
            out["roofline"]["graphed"] = {"ms_per_step": res["graphed_ms_per_step"], "achieved": gg, "frac": gg / pk["hbm_gbs"]}
            if "graphed_inplace_ms_per_step" in res:
                g0 = alg_bytes / (res["graphed_inplace_ms_per_step"] * 1e-3) / 1e9
                out["roofline"]["graphed"].update({"inplace_ms_per_step": res["graphed_inplace_ms_per_step"],
                                                   "inplace_frac": g0 / pk["hbm_gbs"]})
        if "cold_ms_per_step" in res:
            cg = alg_bytes / (res["cold_ms_per_step"] * 1e-3) / 1e9
            out["roofline"]["cold_l2"] = {"ms_per_step": res["cold_ms_per_step"], "achieved": cg, "frac": cg / pk["hbm_gbs"]}
    else:
        out["roofline"] = {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                           "frac": achieved / pk["bf16_tflops"], "frac_of_sustained": achieved / pk["bf16_tflops_sustained"],
                           "peak_source": pk["source"] + " (burst cuBLAS bf16)",
                           # dram__bytes_read.sum + dram__bytes_write.sum of the scan kernel, one launch of this
                           # workload, read from the committed ncu --set full raw page named in traffic_source
                           "traffic": traffic, "traffic_source": traffic_src,
                           "traffic_unit": "bytes per launch (algorithmic: %.3e)" % alg_bytes,
                           "kernel": "scan_tc_kernel (per GPU; step time includes the row-norm and merge kernels)"}
    out["drift_wait_timeouts"] = int(mcl.set_option(103, 0))
    if world == 1 and rank == 0 and not args.profile:
        q, table, labels = res["inputs"]
        out["gpu_unfused_baseline"] = gpu_unfused_baseline(w, q, table, labels, max(2, min(5, args.steps // 10)), device)
        out["gpu_unfused_baseline"]["speedup_fused"] = res["value"] / out["gpu_unfused_baseline"]["value"]
        out["row_kernels"] = row_kernel_rooflines(w, table, device, pk)
    release(res)
    if world == 1 and rank == 0 and not args.profile:
        torch.set_num_threads(os.cpu_count() or 1)
        n = cpu_sample_size(w)
        cpu_reference_step(w, min(n, 64))                 # prepares (and caches) the fp32 table
        t = cpu_reference_step(w, n)
        out["cpu_baseline"] = {"value": n / t, "unit": "queries/s", "cores": torch.get_num_threads(),
                               "kind": "port",
                               "sample": f"{n} of {w['Q']} queries, one step, full table; " + CPU_SAMPLE_NOTE}
        cpu_reference_step.__dict__.get("tables", {}).clear()
        if not args.no_sweep:
            sweep = []
            for name in ("c1", "c2", "gemma3_head", "gemma3_eval", "c4", "c5"):
                if name == args.workload:
                    continue
                try:
                    r = run_workload(name, max(3, args.steps // 4), 3, 1, 0, device, with_e2e=False)
                    ww = WORKLOADS[name]
                    hbm = r["alg_bytes"] / (r["ms_per_step"] * 1e-3) / 1e9
                    entry = {"workload": name, "Q": ww["Q"], "V": ww["V"], "D": ww["D"], "k": ww["k"],
                             "value": r["value"], "ms_per_step": r["ms_per_step"], "tflops": r["tflops"],
                             "tensor_frac": r["tflops"] / pk["bf16_tflops"], "hbm_gbs": hbm,
                             "hbm_frac": hbm / pk["hbm_gbs"], "gpu_launches": r["gpu_launches"],
                             "parity_check": r.get("parity_check")}
                    entry["l2"] = ("inputs fit in L2: 'ms_per_step' is warm, 'cold_ms_per_step' after a 256 MB flush"
                                   if "cold_ms_per_step" in r else "inputs larger than L2")
                    if "cold_ms_per_step" in r:
                        ch = r["alg_bytes"] / (r["cold_ms_per_step"] * 1e-3) / 1e9
                        entry.update({"cold_ms_per_step": r["cold_ms_per_step"], "cold_hbm_gbs": ch,
                                      "cold_hbm_frac": ch / pk["hbm_gbs"]})
                    if "graphed_ms_per_step" in r:      # small batches: CUDA-graph replay of the same step
                        gh = r["alg_bytes"] / (r["graphed_ms_per_step"] * 1e-3) / 1e9
                        entry["graphed"] = {"ms_per_step": r["graphed_ms_per_step"],
                                            "value": ww["Q"] / (r["graphed_ms_per_step"] * 1e-3),
                                            "hbm_gbs": gh, "hbm_frac": gh / pk["hbm_gbs"]}
                        if "graphed_inplace_ms_per_step" in r:    # batch written into the graph's buffer: no input copy
                            g0 = r["alg_bytes"] / (r["graphed_inplace_ms_per_step"] * 1e-3) / 1e9
                            entry["graphed"]["inplace_ms_per_step"] = r["graphed_inplace_ms_per_step"]
                            entry["graphed"]["inplace_hbm_frac"] = g0 / pk["hbm_gbs"]
                    if name in ("c2", "gemma3_head"):
                        qq, tt, ll = r["inputs"]
                        entry["gpu_unfused_baseline"] = gpu_unfused_baseline(ww, qq, tt, ll, 3, device)
                    release(r)
                    if name == "c1":
                        entry["cpu_literal_reference"] = literal_pair_loop_baseline()
                    sweep.append(entry)
                except Exception as e:   # a secondary workload must not void the headline
                    sweep.append({"workload": name, "error": f"{type(e).__name__}: {e}"[:200]})
                    torch.cuda.empty_cache()
            out["sweep"] = sweep
            try:
                out["backward"] = [
                    backward_bench(32768, 1048576, 1024, 32768, device, pk),      # configs[4]: every row labelled
                    backward_bench(1672, 262235, 1152, 24, device, pk),           # the reference's LM head: 3 labels per sample
                    # ... as the drop-in runs it (shims/mllm.py rows="labelled"): only the 48 positions the loss or
                    # the accuracy reads enter the forward (one row block: the panel scan)
                    backward_bench(48, 262235, 1152, 24, device, pk, steps=10),
                ]
            except Exception as e:
                out["backward"] = {"error": f"{type(e).__name__}: {e}"[:300]}
                torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(out)
        pc = out.get("parity_check")
        bad = [pc] if pc and not pc.get("ok", True) else []
        bad += [e["parity_check"] for e in out.get("sweep", []) if e.get("parity_check") and not e["parity_check"].get("ok", True)]
        if bad:
            print("parity_check FAILED: " + json.dumps(bad), file=sys.stderr)
            sys.exit(1)


if __name__ == "__main__":
    main()
