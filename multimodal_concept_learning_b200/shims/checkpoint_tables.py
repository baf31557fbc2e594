"""Shim for the table producer of the analysis scripts:
``load_token_embeddings`` (``src/multimodal/token_embedding_analysis.py:53-124`` and its twin
``token_embedding_analysis_imagenet.py:180-232``).

The reference instantiates the whole MLLM (``from_pretrained`` of Gemma-3 + ViT) and
``load_state_dict``s every ``epoch_*_model.pt`` just to read ONE tensor, the LM input-embedding
table.  Here the table is pulled straight out of each checkpoint's state dict (the on-disk format,
``torch.save(model.state_dict())``, is unchanged), optionally placed on the GPU, and its inverse
row norms -- the only table-dependent quantity the cosine scan needs -- are computed once per
epoch and cached, so a scan of every epoch costs no extra pass over any table."""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch

from .. import ops

# state-dict key of ``MLLM.language_model.get_input_embeddings().weight`` (mllm.py:107);
# Gemma-3 ties lm_head to it (modeling_gemma3.py:593), so either key names the same tensor.
TABLE_KEY_SUFFIXES = ("language_model.model.embed_tokens.weight", "embed_tokens.weight",
                      "language_model.lm_head.weight")


def find_table_key(state_dict) -> str:
    for suffix in TABLE_KEY_SUFFIXES:
        for key in state_dict:
            if key.endswith(suffix):
                return key
    raise KeyError("no input-embedding table in the checkpoint (looked for *%s)" % TABLE_KEY_SUFFIXES[0])


def epoch_checkpoints(models_dir: str) -> Dict[str, str]:
    """The files the reference reads, in its order: initial_model.pt, then epoch_N_model.pt by N
    (``best_model.pt`` is skipped, as in the reference :108)."""
    out: Dict[str, str] = {}
    p = os.path.join(models_dir, "initial_model.pt")
    if os.path.exists(p):
        out["initial"] = p
    files = [f for f in os.listdir(models_dir) if f.startswith("epoch_") and f.endswith("_model.pt")]
    files.sort(key=lambda x: int(x.split("_")[1]))
    for f in files:
        out[f"epoch_{f.split('_')[1]}"] = os.path.join(models_dir, f)
    return out


def load_table(path: str) -> torch.Tensor:
    """One [V, D] table from one checkpoint file, without building any model."""
    try:
        sd = torch.load(path, map_location="cpu", mmap=True, weights_only=True)
    except (RuntimeError, ValueError, TypeError):   # legacy (non-zip) checkpoints cannot be mmapped
        sd = torch.load(path, map_location="cpu", weights_only=True)
    return sd[find_table_key(sd)].detach().clone()   # clone: drop the mmap / the rest of the file


@dataclass
class EpochTables:
    """``embeddings_by_epoch`` (the reference's dict) plus the per-epoch cached inverse norms."""
    tables: Dict[str, torch.Tensor] = field(default_factory=dict)
    inv_norms: Dict[str, torch.Tensor] = field(default_factory=dict)

    _graphs: Dict[tuple, tuple] = field(default_factory=dict, repr=False)

    def scan(self, epoch: str, q: torch.Tensor, k: int, **kw) -> ops.ScanOutput:
        """Cosine top-k / LSE of ``q`` against the table of ``epoch`` using the cached norms."""
        return ops.concept_scan(q, self.tables[epoch], k, inv_norm_t=self.inv_norms[epoch], **kw)

    def scan_all_epochs(self, q: torch.Tensor, k: int, scale: float = 1.0) -> Dict[str, ops.ScanOutput]:
        """The same concept queries against EVERY epoch's table with one launch from the host: the
        per-epoch scans (tables on the GPU, cached inverse norms, no extra pass over any table) are
        captured once per query shape in a CUDA graph and replayed -- the analysis scripts' loop
        over ``embeddings_by_epoch`` (token_embedding_analysis.py:97-121 loads them, :648-660 walks
        them) becomes one ``cudaGraphLaunch`` for all epochs.  The returned outputs alias buffers
        owned by the graph: valid until the next call with the same query shape."""
        if not self.tables or any(not t.is_cuda for t in self.tables.values()):
            raise RuntimeError("scan_all_epochs needs the tables on the GPU: load_embeddings_by_epoch(..., device='cuda')")
        dev = next(iter(self.tables.values())).device
        key = (tuple(q.shape), q.dtype, int(k), float(scale))
        if key not in self._graphs:
            q_static = torch.zeros(q.shape, dtype=q.dtype, device=dev)

            def run():
                return {name: ops.concept_scan(q_static, t, k, scale=scale, inv_norm_t=self.inv_norms[name])
                        for name, t in self.tables.items()}
            with torch.cuda.device(dev):
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):          # one-time work (attributes, tensor maps, plans) outside the capture
                    run()
                torch.cuda.current_stream(dev).wait_stream(side)
                torch.cuda.synchronize(dev)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    outs = run()
            self._graphs[key] = (graph, q_static, outs)
        graph, q_static, outs = self._graphs[key]
        q_static.copy_(q, non_blocking=True)
        graph.replay()
        return outs


def load_embeddings_by_epoch(results_dir: str, device: Optional[str] = None,
                             verbose: bool = True) -> EpochTables:
    """``embeddings_by_epoch`` exactly as ``load_token_embeddings`` builds it (same keys, same
    tensors, same dtype).  ``device=None`` keeps the tables on the CPU like the reference;
    ``device="cuda"`` uploads each once and caches 1/||row|| for the scans."""
    models_dir = os.path.join(results_dir, "models")
    if not os.path.isdir(models_dir):
        raise FileNotFoundError(f"models directory not found at {models_dir}")
    out = EpochTables()
    for name, path in epoch_checkpoints(models_dir).items():
        t = load_table(path)
        if device is not None:
            t = t.to(device)
            if t.dtype in (torch.bfloat16, torch.float32):
                out.inv_norms[name] = ops.row_inv_norm(t)
        out.tables[name] = t
        if verbose:
            label = name.split("_")[1] if name.startswith("epoch_") else name
            print(f"Loaded {'epoch ' + label if name != 'initial' else 'initial model'} embeddings: {t.shape}")
    if verbose:
        print(f"Total loaded {len(out.tables)} embedding matrices")
    return out
