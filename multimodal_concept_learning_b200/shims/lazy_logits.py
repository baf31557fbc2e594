"""``LazyLogits``: what ``outputs.logits`` is under the fused forwards.

The reference reads the logits of its model outputs in exactly three ways:
``torch.argmax(logits, dim=-1)`` (``multimodal_training.py:276``),
``torch.max(outputs.logits.data, 1)`` (``vision_training.py:132,153,228``) and
``criterion(outputs.logits, labels)`` (``vision_training.py:116,149,223``).  A ``LazyLogits`` holds the
operands of the head -- hidden states / features, the weight table, the optional bias and soft-cap
-- and answers those three from the fused scan (k = 1 epilogue: running argmax + online
log-sum-exp, no ``[rows, V]`` matrix).  Anything else (``logits[0]``, ``.float()``, arithmetic, any
other torch function) transparently materialises ``features @ weight.T (+ bias)`` once, in the
dtype the reference's ``nn.Linear`` would have produced, and carries on with the real tensor -- so
unpatched call sites keep working, merely without the saving."""
from __future__ import annotations

from typing import Optional

import torch

from .. import ops
from ._common import compute_device, to_kernel_dtype

_REDUCING = {"argmax", "max"}


def _last_dim(dim, ndim) -> bool:
    return dim is not None and (dim == -1 or dim == ndim - 1)


class LazyLogits:
    def __init__(self, features: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
                 softcap: Optional[float] = None, out_dtype: Optional[torch.dtype] = None):
        self.features, self.weight, self.bias = features, weight, bias
        self.softcap = float(softcap) if softcap else None
        self.out_dtype = out_dtype or features.dtype
        self._dense: Optional[torch.Tensor] = None
        self._top1 = None

    # ---- tensor-like surface --------------------------------------------------------------
    @property
    def shape(self):
        return torch.Size((*self.features.shape[:-1], self.weight.shape[0]))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return self.features.dim()

    @property
    def dtype(self):
        return self.out_dtype

    @property
    def device(self):
        return self.features.device

    @property
    def data(self):          # `outputs.logits.data` (vision_training.py:132): still lazy
        return self

    def detach(self):
        return self

    # ---- the fused answers ------------------------------------------------------------------
    def _operands(self):
        """(q [rows, D'], table [V, D']) in a kernel dtype; a bias rides along as one extra K
        column (features get a 1, the table gets the bias), padded to keep rows 16-byte aligned."""
        dev = compute_device(self.features, self.weight)
        f = to_kernel_dtype(self.features).to(dev).reshape(-1, self.features.shape[-1])
        w = to_kernel_dtype(self.weight).to(device=dev, dtype=f.dtype)
        if self.bias is not None:
            pad = 8 if f.dtype == torch.bfloat16 else 4
            fe = torch.zeros((f.shape[0], pad), dtype=f.dtype, device=dev)
            fe[:, 0] = 1
            we = torch.zeros((w.shape[0], pad), dtype=w.dtype, device=dev)
            we[:, 0] = self.bias.to(device=dev, dtype=w.dtype)
            f, w = torch.cat([f, fe], 1), torch.cat([w, we], 1)
        return f, w

    def top1(self):
        """(values, indices) of the row-wise maximum, first maximum wins (``torch.max`` / ``argmax``)."""
        if self._top1 is None:
            with torch.no_grad():
                f, w = self._operands()
                out = ops.concept_scan(f.detach(), w.detach(), 1, normalize_q=False, normalize_t=False,
                                       softcap=self.softcap)
            lead = self.features.shape[:-1]
            self._top1 = (out.topk_val[:, 0].reshape(lead).to(self.out_dtype), out.topk_idx[:, 0].reshape(lead))
        return self._top1

    def argmax(self, dim=None, keepdim=False):
        if not _last_dim(dim, self.dim()) or keepdim:
            return self.materialize().argmax(dim=dim, keepdim=keepdim)
        return self.top1()[1]

    def max(self, dim=None, keepdim=False):
        if not _last_dim(dim, self.dim()) or keepdim:
            return self.materialize().max() if dim is None else self.materialize().max(dim, keepdim)
        return torch.return_types.max(self.top1())

    def cross_entropy(self, labels: torch.Tensor, label_smoothing: float = 0.0, ignore_index: int = -100):
        """Mean CE over the rows whose label is not ``ignore_index`` -- differentiable with respect
        to the features, the weight and the bias when they require grad (training loops)."""
        if ignore_index != -100:
            labels = torch.where(labels == ignore_index, torch.full_like(labels, -100), labels)
        f, w = self._operands()
        lab = labels.reshape(-1).to(f.device)
        if torch.is_grad_enabled() and (f.requires_grad or w.requires_grad):
            from ..autograd import fused_cross_entropy
            loss, top1 = fused_cross_entropy(f, w, lab, label_smoothing=label_smoothing, softcap=self.softcap)
            return loss
        out = ops.concept_scan(f.detach(), w.detach(), 1, normalize_q=False, normalize_t=False, labels=lab,
                               label_smoothing=label_smoothing, softcap=self.softcap)
        if self._top1 is None:
            lead = self.features.shape[:-1]
            self._top1 = (out.topk_val[:, 0].reshape(lead).to(self.out_dtype), out.topk_idx[:, 0].reshape(lead))
        return out.loss

    # ---- everything else: the real tensor ------------------------------------------------------
    def materialize(self) -> torch.Tensor:
        if self._dense is None:
            z = torch.nn.functional.linear(self.features, self.weight.to(self.features.dtype),
                                           None if self.bias is None else self.bias.to(self.features.dtype))
            if self.softcap:
                z = torch.tanh(z / self.softcap) * self.softcap
            self._dense = z.to(self.out_dtype)
        return self._dense

    def __getattr__(self, name):          # only reached for attributes not defined above
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    def __getitem__(self, item):
        return self.materialize()[item]

    def __len__(self):
        return self.shape[0]

    def __repr__(self):
        return f"LazyLogits(shape={tuple(self.shape)}, dtype={self.out_dtype}, materialized={self._dense is not None})"

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = getattr(func, "__name__", "")
        if name in _REDUCING and args and isinstance(args[0], LazyLogits):
            dim = kwargs.get("dim", args[1] if len(args) > 1 else None)
            keepdim = kwargs.get("keepdim", args[2] if len(args) > 2 else False)
            if isinstance(dim, int):
                return getattr(args[0], name)(dim, keepdim)
        if name == "cross_entropy" and args and isinstance(args[0], LazyLogits) and args[0].dim() == 2 \
                and kwargs.get("weight") is None and kwargs.get("reduction", "mean") == "mean":
            return args[0].cross_entropy(args[1] if len(args) > 1 else kwargs["target"],
                                         label_smoothing=kwargs.get("label_smoothing", 0.0),
                                         ignore_index=kwargs.get("ignore_index", -100))

        def dense(x):
            if isinstance(x, LazyLogits):
                return x.materialize()
            if isinstance(x, (list, tuple)):
                return type(x)(dense(y) for y in x)
            return x
        return func(*dense(args), **{k: dense(v) for k, v in kwargs.items()})


def _forward_operator(name):
    def op(self, *args):
        return getattr(self.materialize(), name)(*args)
    op.__name__ = name
    return op


for _name in ("__add__", "__radd__", "__sub__", "__rsub__", "__mul__", "__rmul__", "__truediv__",
              "__rtruediv__", "__neg__", "__pow__", "__matmul__", "__lt__", "__le__", "__gt__", "__ge__",
              "__iter__", "__float__", "__bool__"):
    setattr(LazyLogits, _name, _forward_operator(_name))
