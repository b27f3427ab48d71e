"""Layer-level operand-rounding mode of the oracle (TEST INFRASTRUCTURE ONLY).

``set_round_operands(module, True)`` switches every oracle LSH layer, FeedForward and cross-attention under ``module``
to CPU ``autograd.Function``s that compute the reference arithmetic (same formulas as the exact path, checked against
it with the roundings off in tests/test_oracle.py) but round to bf16 exactly where the CUDA path stores a bf16 operand:

  LSH layer (rp R1-R11 / hf:405-670)   xn = bf16(LayerNorm out), qk|v = bf16(xn W^T), attention core as
                                       ``oracle.lsh_rounded`` (P, per-round o, merged out; Pt, dS', per-round partials,
                                       dqk|dv), dyb = bf16(dy), dout = bf16(dyb W_out); every contraction accumulates in
                                       fp32 and weight / input gradients are fp32
  FeedForward (ref:reformer_tts/model/modules.py:195-207)
                                       xn = bf16(LayerNorm out), h = bf16(relu(xn W1^T + b1)), dyb = bf16(dy),
                                       dh = bf16((dyb W2) * 1[h > 0])
  cross-attention (ref:reformer_tts/model/reformer.py:161-186)
                                       xn, mem -> bf16, q | k | v = bf16(projection + bias), P -> bf16 (normaliser: fp32 sum
                                       of the unrounded terms), o -> bf16; backward P, dS -> bf16, dq | dk | dv -> bf16

The LayerNorm itself stays the stock fp32 ``nn.LayerNorm`` of the oracle model: the product's LayerNorm kernels keep fp32
statistics and return fp32 gradients, only their OUTPUT is stored in bf16 - which is the first rounding listed above.
Master weights are rounded to bf16 on use (the product multiplies bf16 copies of its fp32 master weights).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

from . import lsh_core, lsh_rounded
from .lsh_rounded import bf16r


def set_round_operands(module: nn.Module, on: bool = True) -> nn.Module:
    for m in module.modules():
        if hasattr(m, "round_operands"):
            m.round_operands = on
    return module


# ---------------------------------------------------------------------------------------------------------------------
class LSHLayerFn(torch.autograd.Function):
    """x_normed fp32 [B,T,D] -> y fp32 [B,T,D]: QK|V projection, (given) buckets, sort, chunked attention, merge, optional
    output projection - the stages of reformer_tts_b200.lsh_attention._LSHAttentionFn after its LayerNorm."""

    @staticmethod
    def forward(ctx, x, w_qk, w_v, w_out, b_out, buckets, mask_bh, cfg):
        b, t, d = x.shape
        h, r, bucket, spec = cfg["heads"], cfg["n_hashes"], cfg["bucket_size"], cfg["spec"]
        dh = d // h
        xn = bf16r(x.detach())
        wqk, wv = bf16r(w_qk.detach()), bf16r(w_v.detach())
        to_bh = lambda a: a.view(b, t, h, dh).transpose(1, 2).reshape(b * h, t, dh)
        qk = to_bh(bf16r(xn @ wqk.t()))
        v = to_bh(bf16r(xn @ wv.t()))
        if buckets is None:
            rot = cfg["rot"]
            rot_bh = rot if rot.shape[0] == 1 else rot[None].expand(b, -1, -1, -1, -1).reshape(b * h, *rot.shape[1:])
            buckets = lsh_core.hash_buckets(qk, rot_bh, r, cfg["n_buckets"], mask_bh if cfg.get("pad_bucket") else None)
        sticker, undo = lsh_core.sort_buckets(buckets, t)
        res = lsh_rounded.forward(qk, v, sticker, undo, bucket, r, spec, mask_bh, round_operands=True)
        out = res["out"].view(b, h, t, dh).transpose(1, 2).reshape(b, t, d)
        if w_out is not None:
            y = out @ bf16r(w_out.detach()).t() + b_out.detach()
        else:
            y = out.clone()
        ctx.cfg, ctx.mask_bh = cfg, mask_bh
        ctx.has_out = w_out is not None
        ctx.save_for_backward(xn, wqk, wv, w_out, qk, v, sticker, undo, res["out"], res["lse"], out)
        cfg["_last"] = dict(res, qk=qk, v=v, buckets=buckets, sticker=sticker, undo=undo)
        return y

    @staticmethod
    def backward(ctx, dy):
        xn, wqk, wv, w_out, qk, v, sticker, undo, out_bh, lse, out = ctx.saved_tensors
        cfg = ctx.cfg
        b, t, d = xn.shape
        h, r, bucket, spec = cfg["heads"], cfg["n_hashes"], cfg["bucket_size"], cfg["spec"]
        dh = d // h
        to_bh = lambda a: a.view(b, t, h, dh).transpose(1, 2).reshape(b * h, t, dh)
        from_bh = lambda a: a.view(b, h, t, dh).transpose(1, 2).reshape(b, t, d)
        g_wout = g_bout = None
        if ctx.has_out:
            dyb = bf16r(dy)
            g_bout = dy.reshape(-1, d).sum(0)
            g_wout = dyb.reshape(-1, d).t() @ out.reshape(-1, d)
            dout = bf16r(dyb @ bf16r(w_out.detach()))
        else:
            dout = bf16r(dy)
        dqk, dv = lsh_rounded.backward(qk, v, sticker, undo, bucket, r, spec, ctx.mask_bh, to_bh(dout), out_bh, lse, round_operands=True)
        dqk, dv = from_bh(dqk).reshape(-1, d), from_bh(dv).reshape(-1, d)
        x2 = xn.reshape(-1, d)
        g_wqk, g_wv = dqk.t() @ x2, dv.t() @ x2
        dxn = (dqk @ wqk + dv @ wv).view(b, t, d)
        return dxn, g_wqk, g_wv, g_wout, g_bout, None, None, None


def lsh_layer(x, w_qk, w_v, w_out, b_out, buckets, mask, heads, bucket_size, n_hashes, spec, rot=None, n_buckets=None,
              pad_bucket=False):
    """Returns (y, stages) - ``stages`` holds qk, v, buckets, sticker, undo, o_rounds, lse_rounds, out, lse of this call."""
    b, t, _ = x.shape
    mask_bh = None if mask is None else mask.bool()[:, None, :].expand(b, heads, t).reshape(b * heads, t)
    cfg = dict(heads=heads, bucket_size=bucket_size, n_hashes=n_hashes, spec=spec, rot=rot, n_buckets=n_buckets, pad_bucket=pad_bucket)
    if buckets is not None:
        buckets = buckets.reshape(b * heads, -1).long()
    y = LSHLayerFn.apply(x, w_qk, w_v, w_out, b_out, buckets, mask_bh, cfg)
    return y, cfg.pop("_last")


# ---------------------------------------------------------------------------------------------------------------------
class FeedForwardFn(torch.autograd.Function):
    """x_normed fp32 -> Linear + ReLU + Linear with the roundings of reformer_tts_b200.feed_forward._LNFeedForwardFn."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        shape = x.shape
        xn = bf16r(x.detach()).reshape(-1, shape[-1])
        w1r, w2r = bf16r(w1.detach()), bf16r(w2.detach())
        hid = bf16r(torch.relu(xn @ w1r.t() + b1.detach()))
        y = hid @ w2r.t() + b2.detach()
        ctx.save_for_backward(xn, hid, w1r, w2r)
        return y.view(shape)

    @staticmethod
    def backward(ctx, dy):
        xn, hid, w1r, w2r = ctx.saved_tensors
        d = xn.shape[1]
        dy2 = dy.reshape(-1, d)
        dyb = bf16r(dy2)
        g_b2 = dy2.sum(0)
        g_w2 = dyb.t() @ hid
        dh32 = (dyb @ w2r) * (hid > 0)
        g_b1 = dh32.sum(0)
        dh = bf16r(dh32)
        g_w1 = dh.t() @ xn
        dxn = dh @ w1r
        return dxn.view(dy.shape), g_w1, g_b1, g_w2, g_b2


# ---------------------------------------------------------------------------------------------------------------------
class CrossAttentionFn(torch.autograd.Function):
    """x_normed fp32 [B,T,D], memory fp32 [B,S,D] -> y fp32: nn.MultiheadAttention arithmetic (no dropout) with the roundings
    of reformer_tts_b200.cross_attention._CrossAttentionFn."""

    @staticmethod
    def forward(ctx, x, memory, w_in, b_in, w_out, b_out, key_padding_mask, heads):
        b, t, d = x.shape
        s = memory.shape[1]
        dh = d // heads
        xn, memb = bf16r(x.detach()), bf16r(memory.detach())
        w = bf16r(w_in.detach())
        bi = b_in.detach()
        q = bf16r(xn @ w[:d].t() + bi[:d])
        k = bf16r(memb @ w[d:2 * d].t() + bi[d:2 * d])
        v = bf16r(memb @ w[2 * d:].t() + bi[2 * d:])
        split = lambda a, n: a.view(b, n, heads, dh).transpose(1, 2)              # [B,H,n,dh]
        ql, kl, vl = split(q, t), split(k, s), split(v, s)
        sc = (ql @ kl.transpose(-1, -2)) / math.sqrt(dh)
        if key_padding_mask is not None:
            sc = sc.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
        m = sc.max(dim=-1, keepdim=True).values
        p = torch.exp(sc - m)
        l = p.sum(dim=-1, keepdim=True)
        o = bf16r((bf16r(p) @ vl) / l)
        lse = (m + torch.log(l)).squeeze(-1)
        o2 = o.transpose(1, 2).reshape(b, t, d)
        wo = bf16r(w_out.detach())
        y = o2 @ wo.t() + b_out.detach()
        ctx.heads, ctx.kpm = heads, key_padding_mask
        ctx.save_for_backward(xn, memb, w, wo, ql, kl, vl, o, lse, o2)
        return y

    @staticmethod
    def backward(ctx, dy):
        xn, memb, w, wo, ql, kl, vl, o, lse, o2 = ctx.saved_tensors
        heads = ctx.heads
        b, t, d = xn.shape
        s = memb.shape[1]
        dh = d // heads
        dyb = bf16r(dy)
        g_bo = dy.reshape(-1, d).sum(0)
        g_wo = dyb.reshape(-1, d).t() @ o2.reshape(-1, d)
        do = bf16r(dyb @ wo).view(b, t, heads, dh).transpose(1, 2)
        sc = (ql @ kl.transpose(-1, -2)) / math.sqrt(dh)
        if ctx.kpm is not None:
            sc = sc.masked_fill(ctx.kpm[:, None, None, :], float("-inf"))
        p = torch.exp(sc - lse[..., None])
        delta = (do * o).sum(dim=-1, keepdim=True)
        dp = do @ vl.transpose(-1, -2)
        ds = p * (dp - delta) / math.sqrt(dh)
        pr, dsr = bf16r(p), bf16r(ds)
        merge = lambda a, n: a.transpose(1, 2).reshape(b * n, d)
        dq = merge(bf16r(dsr @ kl), t)
        dk = merge(bf16r(dsr.transpose(-1, -2) @ ql), s)
        dv = merge(bf16r(pr.transpose(-1, -2) @ do), s)
        x2, m2 = xn.reshape(-1, d), memb.reshape(-1, d)
        g_win = torch.cat([dq.t() @ x2, dk.t() @ m2, dv.t() @ m2], dim=0)
        g_bin = torch.cat([dq.sum(0), dk.sum(0), dv.sum(0)])
        dxn = (dq @ w[:d]).view(b, t, d)
        dmem = (dk @ w[d:2 * d] + dv @ w[2 * d:]).view(b, s, d)
        return dxn, dmem, g_win, g_bin, g_wo, g_bo, None, None
