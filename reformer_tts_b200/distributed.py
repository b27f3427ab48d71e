"""Batch-sharded data parallelism for the ReformerTTS step: one process per GPU, model replicated, gradients
averaged with NCCL all-reduce over NVLink 5 / NVSwitch (the reference is single-GPU; this is new).

Gradient storage.  ``GradientBuckets`` owns ONE flat fp32 buffer; every parameter's ``.grad`` is a view into it, laid out
bucket by bucket: one bucket per reversible block (in the order the reversible backward finishes them) and one for
everything outside the reversible stacks.  Autograd accumulates into the views in place, so a bucket is ready for the
collective the moment its block's backward returns - no ``cat`` into a staging buffer and no copy back - and global-norm
clipping / zeroing are one kernel over the flat buffer.

Overlap.  The reversible backward finishes a block's parameter gradients long before the step ends
(ref:reformer_tts/model/reversible.py:127-128 is the point), so ``GradientAverager`` hooks
``ReversibleSequence.on_block_done`` and issues one asynchronous all-reduce per bucket (NCCL: on the process group's own
stream, ordered after the producing kernels by an event) while the next block recomputes; ``finish()`` issues the last bucket
and makes the compute stream wait for all of them.  Everything is stream / event ordered, so the same code is captured into
the training step's CUDA graph (``reformer_tts_b200.training.TrainStep``) with the collectives as parallel graph branches.
Works with any ``torch.distributed`` backend (``gloo`` on CPU for the tests)."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist
from torch import nn

from .model.reversible import ReversibleSequence
from .residual import register_grad_storage

_ALIGN = 64      # elements (256 B): every bucket starts on a 256-byte boundary


class GradientBuckets:
    """One flat fp32 gradient buffer; ``p.grad`` of every trainable parameter is a view into it."""

    def __init__(self, model: nn.Module):
        params = [p for p in model.parameters() if p.requires_grad]
        if not params:
            raise ValueError("GradientBuckets: the model has no trainable parameters")
        if any(p.dtype != torch.float32 for p in params):
            raise ValueError("GradientBuckets: master weights (and their gradients) are fp32")
        self.device = params[0].device
        seen = set()
        groups: List[List[nn.Parameter]] = []
        self._block_bucket: Dict[int, int] = {}      # id(block) -> bucket index
        for seq in [m for m in model.modules() if isinstance(m, ReversibleSequence)]:
            for block in seq.blocks:
                own = [p for p in block.parameters() if p.requires_grad and id(p) not in seen]
                seen.update(id(p) for p in own)
                if own:
                    self._block_bucket[id(block)] = len(groups)
                    groups.append(own)
        # Everything outside the reversible stacks, one bucket per sub-module (module path, two components deep: `postnet.layers`, `dec.prenet`,
        # `enc.prenet`, `dec.mel_linear`, ...) in parameter order: the modules behind the decoder finish their backward first and the
        # ones in front of the encoder last, so one bucket for all of them could only be reduced after the whole backward - 45 % of the
        # gradient bytes with nothing left to overlap them with.  `rest_buckets` lists them; a bucket is reduced as soon as every one of
        # its parameters has accumulated its gradient (GradientAverager, post-accumulate hooks).
        names = {id(p): n for n, p in model.named_parameters()}
        self.rest_buckets: List[int] = []
        last_key = None
        for p in params:
            if id(p) in seen:
                continue
            key = ".".join(names.get(id(p), "").split(".")[:-1][:2])      # the owning module's path, two components deep
            if key != last_key or not self.rest_buckets:
                self.rest_buckets.append(len(groups))
                groups.append([])
                last_key = key
            groups[-1].append(p)
        self.rest_bucket: Optional[int] = self.rest_buckets[-1] if self.rest_buckets else None
        self.groups = groups
        bounds, total = [], 0
        for g in groups:
            start = total
            for p in g:
                total += p.numel()
            total = (total + _ALIGN - 1) // _ALIGN * _ALIGN
            bounds.append((start, total))
        self.flat = torch.zeros(total, dtype=torch.float32, device=self.device)
        register_grad_storage(self.flat)      # the fused layers' weight-gradient kernels accumulate into it directly (residual.grad_sink)
        self.bounds = bounds
        for g, (start, _) in zip(groups, bounds):
            off = start
            for p in g:
                p.grad = self.flat[off:off + p.numel()].view_as(p)
                off += p.numel()

    def bucket(self, index: int) -> torch.Tensor:
        start, end = self.bounds[index]
        return self.flat[start:end]

    def bucket_of_block(self, block) -> Optional[int]:
        return self._block_bucket.get(id(block))

    def zero(self) -> None:
        self.flat.zero_()

    def attached(self) -> bool:
        """True while every parameter's ``.grad`` still is its view (``zero_grad(set_to_none=True)`` would detach them)."""
        base = self.flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for g in self.groups for p in g)


class GradientAverager:
    """Averages the gradients of a ``GradientBuckets`` over the ranks; per-bucket and overlapped with the reversible
    backward when ``overlap`` (the default).  ``sync_enabled = False`` turns a backward pass into a local accumulation
    (non-boundary micro-batches of gradient accumulation, ref:reformer_tts/training/train.py:77-89)."""

    def __init__(self, model: nn.Module, overlap: bool = True, buckets: Optional[GradientBuckets] = None):
        self.model = model
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.buckets = buckets if buckets is not None else GradientBuckets(model)
        self.overlap = overlap and self.world > 1
        self.sync_enabled = True
        self._works = []
        self._done = set()
        # NCCL averages inside the collective; gloo (CPU tests) sums and the flat buffer is divided once
        self._avg_in_collective = self.world > 1 and dist.get_backend() == "nccl"
        for seq in [m for m in model.modules() if isinstance(m, ReversibleSequence)]:
            seq.on_block_done = self._on_block_done
        # the buckets outside the reversible stacks: reduced when the last of their parameters has accumulated its gradient
        self._pending = {}
        self._hooks = []
        if self.world > 1:
            for index in self.buckets.rest_buckets:
                group = self.buckets.groups[index]
                for p in group:
                    self._hooks.append(p.register_post_accumulate_grad_hook(lambda _p, index=index, n=len(group): self._on_param_done(index, n)))

    # -- internals -------------------------------------------------------------------------------------------------
    def _reduce(self, index: int, async_op: bool):
        if index in self._done:
            return
        self._done.add(index)
        op = dist.ReduceOp.AVG if self._avg_in_collective else dist.ReduceOp.SUM
        work = dist.all_reduce(self.buckets.bucket(index), op=op, async_op=async_op)
        if async_op:
            self._works.append(work)

    def _on_param_done(self, index: int, n_params: int):
        if not (self.overlap and self.sync_enabled and self.world > 1):
            return
        left = self._pending.get(index, n_params) - 1
        self._pending[index] = left
        if left == 0:
            self._reduce(index, async_op=True)

    def _on_block_done(self, index, block):
        if not (self.overlap and self.sync_enabled and self.world > 1):
            return
        bucket = self.buckets.bucket_of_block(block)
        if bucket is not None:
            self._reduce(bucket, async_op=True)

    def enable_overlap(self):
        self.overlap = self.world > 1

    def disable_overlap(self):
        """One flat all-reduce after backward on the current stream instead of per-bucket collectives during it."""
        self.overlap = False

    # -- API -------------------------------------------------------------------------------------------------------
    def finish(self):
        """Call after ``loss.backward()``: reduces what is left and orders the compute stream after every collective."""
        self._pending.clear()
        if self.world == 1 or not self.sync_enabled:
            self._done.clear()
            return
        if self.overlap:
            for index in range(len(self.buckets.groups)):      # the non-block bucket, and any block the hook did not see
                self._reduce(index, async_op=True)
            for work in self._works:
                work.wait()       # CUDA: the current stream waits for the collective's stream (no host blocking)
            self._works.clear()
        else:
            op = dist.ReduceOp.AVG if self._avg_in_collective else dist.ReduceOp.SUM
            dist.all_reduce(self.buckets.flat, op=op)
        self._done.clear()
        if not self._avg_in_collective:
            self.buckets.flat.div_(self.world)


def shard_batch(batch_size: int, rank: int, world: int):
    """Contiguous slice of a global batch owned by ``rank`` (remainder spread over the first ranks)."""
    base, extra = divmod(batch_size, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)
