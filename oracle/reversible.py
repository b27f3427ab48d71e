"""Restatement of ref:reformer_tts/model/reversible.py (TEST INFRASTRUCTURE ONLY).

PINNED: tests/golden/make_golden.py imported the reference file itself in the build container and committed the
gradients it produces for a small stack; tests/test_oracle.py checks this file against them, and also against plain
autograd through the un-reversed composition (KAT-5: the reference's own IrreversibleBlock, ref:...reversible.py:101-111).
CPU only (the RNG record / replay handles the CPU generator; that is all the oracle needs)."""
from __future__ import annotations

import torch
from torch import nn


class Replayable(nn.Module):
    """ref:...reversible.py:11-41 (Deterministic): remember the RNG state of a run, replay it on demand."""

    def __init__(self, net):
        super().__init__()
        self.net = net
        self._state = None

    def forward(self, *args, record_rng=False, set_rng=False, **kwargs):
        if record_rng:
            self._state = torch.get_rng_state()
        if not set_rng:
            return self.net(*args, **kwargs)
        with torch.random.fork_rng(devices=[], enabled=True):
            torch.set_rng_state(self._state)
            return self.net(*args, **kwargs)


def _rerun(fn, inp, grad, retain_graph=False, **kw):
    with torch.enable_grad():
        inp = inp.detach().requires_grad_(True)
        out = fn(inp, set_rng=True, **kw)
        torch.autograd.backward(out, grad, retain_graph=retain_graph)
    return out.detach(), inp.grad


class RevBlock(nn.Module):          # ref:...reversible.py:46-98
    def __init__(self, f, g):
        super().__init__()
        self.f, self.g = Replayable(f), Replayable(g)

    def forward(self, x, f_args={}, g_args={}):
        x1, x2 = x.chunk(2, dim=2)
        with torch.no_grad():
            y1 = x1 + self.f(x2, record_rng=self.training, **f_args)
            y2 = x2 + self.g(y1, record_rng=self.training, **g_args)
        return torch.cat([y1, y2], dim=2)

    def backward_pass(self, y, dy, f_args={}, g_args={}):
        y1, y2 = y.chunk(2, dim=2)
        dy1, dy2 = dy.chunk(2, dim=2)
        gy1, dgy = _rerun(self.g, y1, dy2, **g_args)
        x2, dx1 = y2 - gy1, dy1 + dgy
        fx2, dfx = _rerun(self.f, x2, dx1, retain_graph=True, **f_args)      # :86
        x1, dx2 = y1 - fx2, dy2 + dfx
        return torch.cat([x1, x2], dim=2), torch.cat([dx1, dx2], dim=2)


class RevHalf(nn.Module):           # ref:...reversible.py:134-170
    def __init__(self, f):
        super().__init__()
        self.f = Replayable(f)

    def forward(self, x, **f_args):
        x1, x2 = x.chunk(2, dim=2)
        with torch.no_grad():
            y1 = x1 + self.f(x2, record_rng=self.training, **f_args)
        return torch.cat([y1, x2], dim=2)

    def backward_pass(self, y, dy, **f_args):
        y1, x2 = y.chunk(2, dim=2)
        dy1, dx2 = dy.chunk(2, dim=2)
        # retain_graph=True as at ref:...reversible.py:158: the cross-attention sub-network closes over the encoder output, so
        # every decoder layer's backward walks the WHOLE encoder graph again (the encoder's reversible backward runs
        # dec-depth times per step in the reference).  Kept, because this file is also the timed CPU baseline.
        fx2, dfx = _rerun(self.f, x2, dy1, retain_graph=True, **f_args)
        return torch.cat([y1 - fx2, x2], dim=2), torch.cat([dy1, dx2 + dfx], dim=2)


class RevSwap(nn.Module):           # ref:...reversible.py:173-191
    def forward(self, x, **kwargs):
        a, b = x.chunk(2, dim=2)
        return torch.cat([b, a], dim=2)

    def backward_pass(self, y, dy, **kwargs):
        (b, a), (db, da) = y.chunk(2, dim=2), dy.chunk(2, dim=2)
        return torch.cat([a, b], dim=2), torch.cat([da, db], dim=2)


class _RevFn(torch.autograd.Function):      # ref:...reversible.py:114-129
    @staticmethod
    def forward(ctx, x, blocks, kwargs_list):
        for block, kw in zip(blocks, kwargs_list):
            x = block(x, **kw)
        ctx.y, ctx.blocks, ctx.kwargs_list = x.detach(), blocks, kwargs_list
        return x

    @staticmethod
    def backward(ctx, dy):
        y = ctx.y
        for block, kw in zip(ctx.blocks[::-1], ctx.kwargs_list[::-1]):
            y, dy = block.backward_pass(y, dy, **kw)
        return dy, None, None


class RevSequence(nn.Module):       # ref:...reversible.py:194-203
    def __init__(self, blocks):
        super().__init__()
        self.blocks = blocks

    def forward(self, x, kwargs_list=None):
        kwargs_list = kwargs_list if kwargs_list is not None else [{}] * len(self.blocks)
        return _RevFn.apply(x, self.blocks, kwargs_list)
