"""Helpers shared by the shims: where the computation runs and how host tensors get there."""
from __future__ import annotations

import torch


def compute_device(*tensors) -> torch.device:
    """The CUDA device of the first CUDA tensor, else the current CUDA device.  The shims
    accept the CPU tensors the reference's loaders produce and move what the kernels need
    to the GPU; there is no CPU arithmetic path."""
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("multimodal_concept_learning_b200 needs a CUDA device (sm_100); "
                           "there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def to_kernel_dtype(t: torch.Tensor) -> torch.Tensor:
    """bf16 stays bf16 (tcgen05 path); everything else is widened to fp32 (check path).
    fp16 -> fp32 is exact, so the vision loop's fp16 autocast loses nothing."""
    if t.dtype in (torch.bfloat16, torch.float32):
        return t
    return t.float()
