"""Batch-sharded data parallelism for the ReformerTTS step: one process per GPU, model replicated, gradients
averaged with NCCL all-reduce over NVLink 5 / NVSwitch (the reference is single-GPU; this is new).

The reversible backward finishes a block's parameter gradients long before the step ends
(ref:reformer_tts/model/reversible.py:127-128 is the point), so ``GradientAverager`` hooks
``ReversibleSequence.on_block_done`` and launches one flat all-reduce per block on a side stream while the next
block recomputes; everything outside the reversible stacks goes in one last bucket.  Works with any
``torch.distributed`` backend (``gloo`` on CPU for the tests)."""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist
from torch import nn

from .model.reversible import ReversibleSequence


class GradientAverager:
    def __init__(self, model: nn.Module, overlap: bool = True):
        self.model = model
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.overlap = overlap and self.world > 1
        self._pending = []          # (work handle or stream event, flat buffer, params)
        self._stream = None
        self._block_params = {}     # id(block) -> params
        self._in_blocks = set()
        for seq in [m for m in model.modules() if isinstance(m, ReversibleSequence)]:
            for block in seq.blocks:
                params = [p for p in block.parameters() if p.requires_grad]
                self._block_params[id(block)] = params
                self._in_blocks.update(id(p) for p in params)
            if self.overlap:
                seq.on_block_done = self._on_block_done
        self._rest = [p for p in model.parameters() if p.requires_grad and id(p) not in self._in_blocks]

    # -- internals -------------------------------------------------------------------------------------------------
    def _launch(self, params: List[torch.Tensor]):
        params = [p for p in params if p.grad is not None]
        if not params or self.world == 1:
            return
        flat = torch.cat([p.grad.reshape(-1) for p in params])
        if not self.overlap:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            self._pending.append((None, flat, params))
            return
        if flat.is_cuda:
            if self._stream is None:
                self._stream = torch.cuda.Stream()
            self._stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._stream):
                work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
            flat.record_stream(self._stream)
        else:
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
        self._pending.append((work, flat, params))

    def _on_block_done(self, index, block):
        if self.overlap:
            self._launch(self._block_params.get(id(block), []))

    def enable_overlap(self):
        self.overlap = self.world > 1

    def disable_overlap(self):
        """One flat all-reduce after backward on the current stream (what a CUDA-graph capture of the step records)."""
        self.overlap = False

    # -- API -------------------------------------------------------------------------------------------------------
    def finish(self):
        """Call after ``loss.backward()``: reduces what is left and writes the averaged gradients back."""
        if self.world == 1:
            return
        if self.overlap:
            self._launch(self._rest)
        else:
            self._launch([p for p in self.model.parameters() if p.requires_grad])
        for work, flat, params in self._pending:
            if work is not None:
                work.wait()
                if flat.is_cuda:
                    torch.cuda.current_stream().wait_stream(self._stream)
            flat.div_(self.world)
            offset = 0
            for p in params:
                n = p.grad.numel()
                p.grad.copy_(flat[offset:offset + n].view_as(p.grad))
                offset += n
        self._pending.clear()


def shard_batch(batch_size: int, rank: int, world: int):
    """Contiguous slice of a global batch owned by ``rank`` (remainder spread over the first ranks)."""
    base, extra = divmod(batch_size, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)
