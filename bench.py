#!/usr/bin/env python3
"""bench.py -- concept queries/sec of the fused similarity scan (top-k=50 + LSE/CE).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3]

One "step" = one pass of the hot path over one batch of synthetic queries: every query row
scored against every table row with top-k and log-sum-exp statistics produced
(`mcl_concept_scan[_sharded]`).  The headline workload is BASELINE.json configs[2]
(Qwen2-VL-7B-scale vocabulary, the shape north_star quotes its target on); with N > 1 the
table is sharded by vocabulary rows across the ranks (strong scaling: total work fixed), each
step ending in one NCCL all-gather + merge.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: Q, V, D, cosine?, scale, labels?, BASELINE.json config index
    "c1": dict(Q=16, V=50257, D=768, normalize=True, scale=1.0, labels=False, cfg=0,
               desc="16 concept embeddings x GPT-2 vocab 50257x768, cosine top-50"),
    "c2": dict(Q=4096, V=49408, D=768, normalize=True, scale=100.0, labels=True, cfg=1,
               desc="4096 queries x CLIP ViT-L/14 text table 49408x768, top-50 + softmax-CE"),
    "c3": dict(Q=8192, V=152064, D=3584, normalize=True, scale=1.0, labels=False, cfg=2,
               desc="8192 multi-token concept embeddings x Qwen2-VL-7B vocab 152064x3584, cosine top-50 + LSE"),
    "c4": dict(Q=65536, V=128256, D=4096, normalize=True, scale=1.0, labels=False, cfg=3,
               desc="65536 queries x Llama-3-8B vocab 128256x4096, top-50 + LSE"),
    "c5": dict(Q=32768, V=1048576, D=1024, normalize=True, scale=100.0, labels=True, cfg=4,
               desc="32768 image embeddings x 1M concept bank 1048576x1024, contrastive logits + CE"),
}
K_TOP = 50
METRIC = "concept queries/sec vs vocab (top-k=50)"


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
    g = torch.Generator(device=device).manual_seed(seed_base + w["cfg"])
    q = torch.randn(w["Q"], w["D"], generator=g, device=device).to(torch.bfloat16)
    labels = torch.randint(0, w["V"], (w["Q"],), generator=g, device=device) if w["labels"] else None
    lo, hi = shard_rows(w["V"], world, rank)
    gt = torch.Generator(device=device).manual_seed(seed_base + 100 * (rank + 1) + w["cfg"])
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
    """Returns (step_fn, close_fn): the public-API call a user makes, inputs resident."""
    import multimodal_concept_learning_b200 as mcl
    if world > 1:
        from multimodal_concept_learning_b200.sharded import ShardedConceptScan
        sc = ShardedConceptScan(table, vocab_total, normalize_t=w["normalize"])
        inv_q = mcl.row_inv_norm(q) if w["normalize"] else None

        def step(qq=q):
            return sc.scan(qq, K_TOP, normalize_q=w["normalize"], scale=w["scale"], labels=labels,
                           inv_norm_q=inv_q if qq is q else None)
        step.scanner, step.inv_t = sc, None
        return step, sc.close
    inv_t = mcl.row_inv_norm(table) if w["normalize"] else None   # cached per table version
    inv_q = mcl.row_inv_norm(q) if w["normalize"] else None

    def step(qq=q):
        return mcl.concept_scan(qq, table, K_TOP, normalize_q=w["normalize"], normalize_t=w["normalize"],
                                scale=w["scale"], labels=labels, inv_norm_t=inv_t,
                                inv_norm_q=inv_q if qq is q else None)
    step.scanner, step.inv_t = None, inv_t
    return step, (lambda: None)


def run_workload(name, steps, warmup, world, rank, device, with_e2e=True, sampler=None):
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
        g = mcl.GraphedConceptScan(table, K_TOP, w["Q"], normalize=w["normalize"], scale=w["scale"],
                                   inv_norm_t=step.inv_t, with_labels=labels is not None)
        gms = time_steps(lambda: g(q, labels), steps, warmup, world, device)
        res["graphed_ms_per_step"] = gms / steps
        del g
    flops = 2.0 * w["Q"] * w["V"] * w["D"]
    res["tflops"] = flops / (ms / steps * 1e-3) / 1e12
    res["alg_bytes"] = 2.0 * (w["V"] * w["D"] + w["Q"] * w["D"]) + w["Q"] * (8 * K_TOP + 16)
    if with_e2e:
        # same metric through the public API with HOST buffers: every step copies the query batch
        # from pinned host memory and returns (top-k values, indices, stats) in host memory.
        # HostQueryPipeline overlaps the copies of neighbouring steps with the scan.
        from multimodal_concept_learning_b200.pipeline import HostQueryPipeline
        q_host = q.cpu().pin_memory()
        pipe = HostQueryPipeline(table, K_TOP, normalize=w["normalize"], scale=w["scale"],
                                 inv_norm_t=step.inv_t, scanner=step.scanner, reuse_host_buffers=True)
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
                      "h2d_bytes_per_step": q_host.numel() * q_host.element_size(),
                      "d2h_bytes_per_step": sum(t.numel() * t.element_size() for t in outs)}
    close()
    del q, table
    torch.cuda.empty_cache()
    return res


def cpu_reference_step(w, n_queries, seed=1234):
    """The reference's own CPU path for this workload -- F.normalize -> @ -> topk -> logsumexp /
    cross_entropy in fp32 (oracle/concept_scan_ref.torch_composition_ref) -- on a bounded
    query sample against the FULL table; returns seconds for one step."""
    from oracle.concept_scan_ref import torch_composition_ref
    g = torch.Generator().manual_seed(seed + w["cfg"])
    q = torch.randn(n_queries, w["D"], generator=g).to(torch.bfloat16)
    cache = cpu_reference_step.__dict__.setdefault("tables", {})
    key = (w["V"], w["D"])
    if key not in cache:
        cache.clear()
        cache[key] = torch.randn(w["V"], w["D"], generator=g).to(torch.bfloat16)
    table = cache[key]
    labels = torch.randint(0, w["V"], (n_queries,), generator=g) if w["labels"] else None
    t0 = time.perf_counter()
    torch_composition_ref(q, table, K_TOP, normalize=w["normalize"], scale=w["scale"], labels=labels)
    return time.perf_counter() - t0


def cpu_sample_size(w):
    # ~10-30 s of CPU work on a typical host: table normalisation dominates for small samples
    return max(16, min(w["Q"], int(2.0e11 / (2.0 * w["V"] * w["D"]))))


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
                    help="kernel-only run for ncu: no sweep, no e2e, no CPU baseline (not a bench value)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3                                   # timing rule: W >= 3
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w = WORKLOADS[args.workload]
    pk = peaks()

    base = {"metric": METRIC, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[{w['cfg']}]: {w['desc']}", "Q": w["Q"],
                       "V": w["V"], "D": w["D"], "k": K_TOP,
                       "parallelism": f"vocab-row sharding x{args.gpus}" if args.gpus > 1 else "single GPU",
                       "l2": "inputs larger than L2 (table {:.2f} GB vs 126 MB)".format(w["V"] * w["D"] * 2 / 1e9)}}

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
                                      "sample": f"{n} of {w['Q']} queries per step against the full table "
                                                "(torch fp32 normalize->matmul->topk->logsumexp/CE)"},
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
    res = run_workload(args.workload, args.steps, args.warmup, world, rank, device, not args.profile, sampler)
    out = dict(base)
    out.update({"value": res["value"], "ms_per_step": res["ms_per_step"], "e2e": res.get("e2e"),
                "gpu_launches": res["gpu_launches"], "clocks": res["clocks"]})
    shard_flops = 2.0 * w["Q"] * w["V"] * w["D"] / world
    achieved = shard_flops / (res["ms_per_step"] * 1e-3) / 1e12
    out["roofline"] = {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                       "frac": achieved / pk["bf16_tflops"], "frac_of_sustained": achieved / pk["bf16_tflops_sustained"],
                       "peak_source": pk["source"] + " (burst cuBLAS bf16)",
                       # dram__bytes_read.sum + dram__bytes_write.sum of scan_tc_kernel<2,0>, one launch of this
                       # workload at N=1, from profiles/r01_c3_scan_tc_v8_ncu_raw.csv (ncu --set full)
                       "traffic": 5.665e9 if (args.workload == "c3" and world == 1) else None,
                       "traffic_unit": "bytes per launch (algorithmic: %.3e)" % (
                           2.0 * (w["V"] / world * w["D"] + w["Q"] * w["D"]) + w["Q"] * (8 * K_TOP + 16)),
                       "kernel": "scan_tc_kernel (per GPU; step time includes the merge kernel)"}
    if world == 1 and rank == 0 and not args.profile:
        torch.set_num_threads(os.cpu_count() or 1)
        n = cpu_sample_size(w)
        t = cpu_reference_step(w, n)
        out["cpu_baseline"] = {"value": n / t, "unit": "queries/s", "cores": torch.get_num_threads(),
                               "kind": "port",
                               "sample": f"{n} of {w['Q']} queries, one step, full table, torch fp32 composition"}
        if not args.no_sweep:
            sweep = []
            for name in ("c1", "c2", "c4", "c5"):
                try:
                    r = run_workload(name, max(3, args.steps // 4), 3, 1, 0, device, with_e2e=False)
                    ww = WORKLOADS[name]
                    hbm = r["alg_bytes"] / (r["ms_per_step"] * 1e-3) / 1e9
                    entry = {"workload": name, "Q": ww["Q"], "V": ww["V"], "D": ww["D"],
                             "value": r["value"], "ms_per_step": r["ms_per_step"], "tflops": r["tflops"],
                             "tensor_frac": r["tflops"] / pk["bf16_tflops"], "hbm_gbs": hbm,
                             "hbm_frac": hbm / pk["hbm_gbs"]}
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
                    sweep.append(entry)
                except Exception as e:   # a secondary workload must not void the headline
                    sweep.append({"workload": name, "error": f"{type(e).__name__}: {e}"[:200]})
            out["sweep"] = sweep
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(out)


if __name__ == "__main__":
    main()
