"""Reversible residual machinery with the interface of ref:reformer_tts/model/reversible.py.

Same algebra (RevNet): only the output of the whole sequence is kept; the backward walks the blocks in reverse,
re-runs each sub-network with its recorded RNG state, back-propagates through that one sub-network and reconstructs
the block input by subtraction.  Differences from the reference that do not change results:

* the two halves travel as separate tensors inside ``ReversibleSequence`` (``forward_halves`` /
  ``backward_halves``); the public ``forward(x)`` / ``backward_pass(y, dy)`` taking concatenated tensors are kept
  for drop-in use but the per-block ``torch.cat`` copies (ref:...reversible.py:60,97-98,146,178-179) are gone and
  ``ReversibleSwap`` is free;
* our LSH-attention and FeedForward sub-networks are single autograd.Functions with hand-written backward, so
  "recompute with grad" = one more run of the fused forward kernels, and ``autograd.backward`` = the fused
  backward kernels;
* the residual stream stays fp32 (SURVEY.md hard parts: a bf16 stream breaks the subtractive reconstruction).
"""
from __future__ import annotations

import torch
from torch import nn
from torch.autograd.function import Function
from torch.utils.checkpoint import get_device_states, set_device_states

from ..residual import GradAccumRequest, ResidualRequest


class Deterministic(nn.Module):
    """Runs ``net``; can record the RNG state before a run and replay it later (ref:...reversible.py:11-41), so the
    rotations drawn by the LSH hash and the dropout masks of the recompute equal those of the forward.

    Two mechanisms give that guarantee:
    * default - the reference's: save the CPU / CUDA generator states in the forward, restore them inside ``fork_rng`` for the
      recompute.  Needs host access to the generator state, which CUDA-graph capture forbids.
    * ``use_private_generators(seed)`` - capture-safe: the module owns two CUDA generators with the same seed; the forward run
      consumes one, the recompute the other (the default generator is pointed at them with ``graphsafe_set_state`` for the
      duration of the run), so both see the same stream without ever reading a generator state.  Used by
      ``reformer_tts_b200.training.TrainStep`` when it captures the step in a CUDA graph.
    """

    def __init__(self, net):
        super().__init__()
        self.net = net
        self.cpu_state = None
        self.cuda_in_fwd = None
        self.gpu_devices = None
        self.gpu_states = None
        self._private = None        # (device index, forward generator, recompute generator)

    def use_private_generators(self, seed: int, device: torch.device):
        fwd = torch.Generator(device=device).manual_seed(seed)
        rec = torch.Generator(device=device).manual_seed(seed)
        self._private = (device.index if device.index is not None else torch.cuda.current_device(), fwd, rec)
        return fwd, rec

    def record_rng(self, *args):
        self.cpu_state = torch.get_rng_state()
        if torch.cuda._initialized:
            self.cuda_in_fwd = True
            self.gpu_devices, self.gpu_states = get_device_states(*args)

    def _run_private(self, generator, *args, **kwargs):
        default = torch.cuda.default_generators[self._private[0]]
        saved = default.graphsafe_get_state()
        default.graphsafe_set_state(generator)
        try:
            return self.net(*args, **kwargs)
        finally:
            default.graphsafe_set_state(saved)

    def forward(self, *args, record_rng=False, set_rng=False, **kwargs):
        if self._private is not None and (record_rng or set_rng):
            if record_rng and not torch.cuda.is_current_stream_capturing():
                # a forward without a matching recompute (no_grad evaluation in train mode, an exception, a skipped step) must not
                # leave the pair out of step for good: the recompute generator restarts where this forward starts.  (Generator
                # state is host-side bookkeeping: no device sync.  Inside a capture every forward has its backward.)
                self._private[2].set_state(self._private[1].get_state())
            return self._run_private(self._private[2] if set_rng else self._private[1], *args, **kwargs)
        if record_rng:
            self.record_rng(*args)
        if not set_rng:
            return self.net(*args, **kwargs)
        devices = self.gpu_devices if self.cuda_in_fwd else []
        with torch.random.fork_rng(devices=devices, enabled=True):
            torch.set_rng_state(self.cpu_state)
            if self.cuda_in_fwd:
                set_device_states(self.gpu_devices, self.gpu_states)
            return self.net(*args, **kwargs)


def _takes_residual(fn) -> bool:
    """True when the sub-network is a (WithNorm / Chunk wrapped) layer of this library whose LAST kernel is a GEMM that can take
    the residual in its epilogue (FeedForward, the cross-attention block).  Anything else is never offered one."""
    net = getattr(fn, "net", fn)
    for _ in range(4):
        if getattr(net, "takes_residual", False):
            return True
        if type(net).__name__ not in ("WithNorm", "Chunk") or not hasattr(net, "fn"):
            return False
        net = net.fn
    return False


def _grad_through(fn, inp, grad_out, y, acc, **kwargs):
    """Re-run ``fn`` on a detached copy of ``inp`` with grad enabled, push ``grad_out`` through it, reconstruct the block
    input ``y - fn(inp)`` and add the input gradient onto ``acc``.  Returns (y - fn(inp), acc + d loss / d inp).  Parameter
    gradients accumulate into ``.grad`` as usual.  Both elementwise passes are offered to the sub-network: a layer ending in
    one of this library's GEMMs writes ``y - f`` from its epilogue (``ResidualRequest``), and one whose backward ends in the
    LayerNorm-backward kernel adds ``acc`` there (``GradAccumRequest``)."""
    with torch.enable_grad():
        leaf = inp.detach().requires_grad_(True)
        if _takes_residual(fn):
            with ResidualRequest(y, "reconstruct") as req:
                out = fn(leaf, set_rng=True, **kwargs)
            fused = req.consumed
        else:
            out, fused = fn(leaf, set_rng=True, **kwargs), False
        if _takes_residual(fn):
            with GradAccumRequest(acc) as greq:
                torch.autograd.backward(out, grad_out)
            acc_fused = greq.consumed
        else:
            torch.autograd.backward(out, grad_out)
            acc_fused = False
    out = out.detach()
    return (out if fused else y - out), (leaf.grad if acc_fused else acc + leaf.grad)


def _residual_forward(fn, x_res, inp, **kwargs):
    """``x_res + fn(inp)`` with the addition offered to the sub-network's last GEMM epilogue."""
    if not _takes_residual(fn):
        return x_res + fn(inp, **kwargs)
    with ResidualRequest(x_res, "add") as req:
        out = fn(inp, **kwargs)
    return out if req.consumed else x_res + out


class ReversibleBlock(nn.Module):
    """y1 = x1 + f(x2); y2 = x2 + g(y1)   (ref:...reversible.py:46-98)."""

    def __init__(self, f, g):
        super().__init__()
        self.f = Deterministic(f)
        self.g = Deterministic(g)

    def forward_halves(self, x1, x2, f_args={}, g_args={}):
        with torch.no_grad():
            y1 = _residual_forward(self.f, x1, x2, record_rng=self.training, **f_args)
            y2 = _residual_forward(self.g, x2, y1, record_rng=self.training, **g_args)
        return y1, y2

    def backward_halves(self, y1, y2, dy1, dy2, f_args={}, g_args={}):
        with torch.no_grad():
            dy1, dy2 = dy1.contiguous(), dy2.contiguous()      # (halves of one tensor at the first block: views with a row stride)
        x2, dx1 = _grad_through(self.g, y1, dy2, y2, dy1, **g_args)         # x2 = y2 - g(y1), dx1 = dy1 + dg
        x1, dx2 = _grad_through(self.f, x2, dx1, y1, dy2, **f_args)         # x1 = y1 - f(x2), dx2 = dy2 + df
        return x1, x2, dx1, dx2

    def forward(self, x, f_args={}, g_args={}):
        x1, x2 = torch.chunk(x, 2, dim=2)
        return torch.cat(self.forward_halves(x1, x2, f_args, g_args), dim=2)

    def backward_pass(self, y, dy, f_args={}, g_args={}):
        y1, y2 = torch.chunk(y, 2, dim=2)
        dy1, dy2 = torch.chunk(dy, 2, dim=2)
        x1, x2, dx1, dx2 = self.backward_halves(y1, y2, dy1, dy2, f_args, g_args)
        return torch.cat([x1, x2], dim=2), torch.cat([dx1, dx2], dim=2)


class ReversibleHalfResidual(nn.Module):
    """y1 = x1 + f(x2); y2 = x2   (ref:...reversible.py:134-170)."""

    def __init__(self, f):
        super().__init__()
        self.f = Deterministic(f)

    def forward_halves(self, x1, x2, **f_args):
        with torch.no_grad():
            y1 = _residual_forward(self.f, x1, x2, record_rng=self.training, **f_args)
        return y1, x2

    def backward_halves(self, y1, x2, dy1, dx2, **f_args):
        with torch.no_grad():
            dx2 = dx2.contiguous()
        x1, dx2 = _grad_through(self.f, x2, dy1, y1, dx2, **f_args)         # x1 = y1 - f(x2), dx2 += df
        return x1, x2, dy1, dx2

    def forward(self, x, **f_args):
        x1, x2 = torch.chunk(x, 2, dim=2)
        return torch.cat(self.forward_halves(x1, x2, **f_args), dim=2)

    def backward_pass(self, y, dy, **f_args):
        y1, x2 = torch.chunk(y, 2, dim=2)
        dy1, dx2 = torch.chunk(dy, 2, dim=2)
        x1, x2, dx1, dx2 = self.backward_halves(y1, x2, dy1, dx2, **f_args)
        return torch.cat([x1, x2], dim=2), torch.cat([dx1, dx2], dim=2)


class ReversibleSwap(nn.Module):
    """(x1, x2) -> (x2, x1)   (ref:...reversible.py:173-191); a no-op on separate halves."""

    def forward_halves(self, x1, x2, **kwargs):
        return x2, x1

    def backward_halves(self, y1, y2, dy1, dy2, **kwargs):
        return y2, y1, dy2, dy1

    def forward(self, x, **kwargs):
        x1, x2 = torch.chunk(x, 2, dim=2)
        return torch.cat([x2, x1], dim=2)

    def backward_pass(self, y, dy, **kwargs):
        x2, x1 = torch.chunk(y, 2, dim=2)
        dx2, dx1 = torch.chunk(dy, 2, dim=2)
        return torch.cat([x1, x2], dim=2), torch.cat([dx1, dx2], dim=2)


def _swap_tensors(obj, mapping):
    """Copy of a (nested) kwargs structure with tensors replaced according to ``mapping`` (id -> tensor)."""
    if isinstance(obj, torch.Tensor):
        return mapping.get(id(obj), obj)
    if isinstance(obj, dict):
        return {k: _swap_tensors(v, mapping) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_swap_tensors(v, mapping) for v in obj)
    return obj


def _differentiable_tensors(obj, found):
    if isinstance(obj, torch.Tensor):
        if obj.requires_grad and all(obj is not t for t in found):
            found.append(obj)
    elif isinstance(obj, dict):
        for v in obj.values():
            _differentiable_tensors(v, found)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            _differentiable_tensors(v, found)


class _ReversibleFunction(Function):
    """ref:...reversible.py:114-129: forward keeps only the final activation; backward reconstructs.

    Tensors inside ``kwargs_list`` that require grad (the encoder output the decoder's cross-attention closes over)
    are passed as explicit inputs: in backward each block sees a detached leaf, the leaf accumulates the gradient of
    every layer, and the sum flows back ONCE.  The reference instead back-propagates through the encoder graph once
    per decoder layer with retain_graph=True (ref:...reversible.py:158); the gradient is the same sum."""

    @staticmethod
    def forward(ctx, x, blocks, kwargs_list, on_block_done, *context):
        x1, x2 = torch.chunk(x, 2, dim=2)
        for block, kwargs in zip(blocks, kwargs_list):
            if hasattr(block, "forward_halves"):
                x1, x2 = block.forward_halves(x1, x2, **kwargs)
            else:       # foreign block with the reference interface only
                x1, x2 = torch.chunk(block(torch.cat([x1, x2], dim=2), **kwargs), 2, dim=2)
        ctx.y1, ctx.y2 = x1.detach(), x2.detach()
        ctx.blocks, ctx.kwargs_list, ctx.on_block_done, ctx.context = blocks, kwargs_list, on_block_done, context
        return torch.cat([x1, x2], dim=2)

    @staticmethod
    def backward(ctx, dy):
        y1, y2 = ctx.y1, ctx.y2
        dy1, dy2 = torch.chunk(dy, 2, dim=2)
        leaves = [t.detach().requires_grad_(True) for t in ctx.context]
        kwargs_list = _swap_tensors(ctx.kwargs_list, {id(t): leaf for t, leaf in zip(ctx.context, leaves)}) if leaves else ctx.kwargs_list
        for index in range(len(ctx.blocks) - 1, -1, -1):
            block, kwargs = ctx.blocks[index], kwargs_list[index]
            if hasattr(block, "backward_halves"):
                y1, y2, dy1, dy2 = block.backward_halves(y1, y2, dy1, dy2, **kwargs)
            else:
                y, d = block.backward_pass(torch.cat([y1, y2], dim=2), torch.cat([dy1, dy2], dim=2), **kwargs)
                (y1, y2), (dy1, dy2) = torch.chunk(y, 2, dim=2), torch.chunk(d, 2, dim=2)
            if ctx.on_block_done is not None:
                ctx.on_block_done(index, block)     # gradients of this block's parameters are final here
        return (torch.cat([dy1, dy2], dim=2), None, None, None, *[leaf.grad for leaf in leaves])


class ReversibleSequence(nn.Module):
    """ref:...reversible.py:194-203.  ``on_block_done(index, block)`` (optional attribute) is called in backward as
    soon as a block's parameter gradients are complete - the hook the data-parallel gradient all-reduce overlaps on."""

    def __init__(self, blocks):
        super().__init__()
        self.blocks = blocks
        self.on_block_done = None

    def forward(self, x, kwargs_list=None, **kwargs):
        blocks = self.blocks
        blocks_kwargs = kwargs_list if kwargs_list is not None else [{}] * len(blocks)
        context = []
        if torch.is_grad_enabled():
            _differentiable_tensors(blocks_kwargs, context)
        return _ReversibleFunction.apply(x, blocks, blocks_kwargs, self.on_block_done, *context)
