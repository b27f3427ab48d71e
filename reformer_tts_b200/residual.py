"""Hand-offs between the reversible blocks / the trainer and the fused layers (no kernels here: importable without the CUDA
library): the reversible residual offered to a sub-network's last GEMM epilogue, the gradient accumulation offered to its
LayerNorm-backward kernel, and the flat gradient buffer the weight-gradient kernels accumulate into directly."""
from __future__ import annotations

from typing import Optional

import torch


class ResidualRequest:
    """A reversible block's offer to the sub-network it is about to run: "your last GEMM may write ``resid + f`` (mode 'add', the
    reversible forward) or ``resid - f`` (mode 'reconstruct', the recompute in the reversible backward) straight from its epilogue".
    A layer whose last kernel is one of this library's GEMMs with an fp32 output takes the offer (``take()``) and returns the
    combined tensor; the block checks ``consumed`` and otherwise adds / subtracts itself.  In 'reconstruct' mode the returned tensor
    is x1 = y1 - f(x2) while its autograd node still maps an incoming gradient as d loss / d f (the block never differentiates
    through the reconstruction)."""
    current: Optional["ResidualRequest"] = None

    def __init__(self, resid: torch.Tensor, mode: str):
        assert mode in ("add", "reconstruct")
        self.resid, self.mode, self.consumed = resid, mode, False

    def __enter__(self):
        self._prev, ResidualRequest.current = ResidualRequest.current, self
        return self

    def __exit__(self, *exc):
        ResidualRequest.current = self._prev
        return False

    @staticmethod
    def take(shape, device):
        """-> (resid as fp32 [rows, cols] or None, subtract?)  Marks the pending request as consumed when it fits."""
        req = ResidualRequest.current
        if req is None or req.consumed:
            return None, False
        r = req.resid
        if r.dtype != torch.float32 or r.device != device or tuple(r.shape) != tuple(shape) or not r.is_contiguous():
            return None, False
        req.consumed = True
        return r.view(-1, shape[-1]), req.mode == "reconstruct"


class GradAccumRequest:
    """The backward-pass counterpart: a reversible block is about to back-propagate through a sub-network and will add the
    resulting input gradient onto ``base`` (``dx2 = dx2 + df``).  A layer whose backward ends in this library's LayerNorm-backward
    kernel takes the request and returns ``base + df`` as its input gradient; the block checks ``consumed``."""
    current: Optional["GradAccumRequest"] = None

    def __init__(self, base: torch.Tensor):
        self.base, self.consumed = base, False

    def __enter__(self):
        self._prev, GradAccumRequest.current = GradAccumRequest.current, self
        return self

    def __exit__(self, *exc):
        GradAccumRequest.current = self._prev
        return False

    @staticmethod
    def take(numel, device):
        req = GradAccumRequest.current
        if req is None or req.consumed:
            return None
        b = req.base
        if b.dtype != torch.float32 or b.device != device or b.numel() != numel or not b.is_contiguous():
            return None
        req.consumed = True
        return b


# Storages of the flat gradient buffers (distributed.GradientBuckets) whose views are the parameters' ``.grad``.
_SINK_STORAGES = set()


def register_grad_storage(flat: torch.Tensor) -> None:
    _SINK_STORAGES.add(flat.untyped_storage().data_ptr())


def grad_sink(*params) -> Optional[torch.Tensor]:
    """The ``.grad`` buffers of ``params`` as ONE fp32 tensor (rows stacked along dim 0) when they are views of a registered flat
    gradient buffer lying back to back in it (GradientBuckets lays a block's parameters out in declaration order), else None.
    A hand-written backward then lets its weight-gradient kernel (split-K ``red.global.add``, column-sum atomics) accumulate
    straight into the buffer and returns ``None`` for those inputs: no zero-filled temporary, no separate ``grad += `` pass."""
    grads = []
    for p in params:
        g = getattr(p, "grad", None)
        if (g is None or g.dtype != torch.float32 or not g.is_contiguous()
                or g.untyped_storage().data_ptr() not in _SINK_STORAGES):
            return None
        grads.append(g)
    first = grads[0]
    if len(grads) == 1:
        return first
    offset = first.storage_offset()
    for g in grads:
        if g.storage_offset() != offset or g.shape[1:] != first.shape[1:] or g.device != first.device:
            return None
        offset += g.numel()
    rows = sum(g.shape[0] for g in grads)
    return torch.as_strided(first, (rows,) + tuple(first.shape[1:]), first.stride(), first.storage_offset())
