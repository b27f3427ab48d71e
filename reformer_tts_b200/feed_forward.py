"""Fused LayerNorm + FeedForward (ref:reformer_tts/model/modules.py:195-207 under ref:reformer_tts/model/reformer.py:25-45).

The reference runs ``Chunk(100, WithNorm(LayerNorm, FeedForward))``: ~90 slices of 3-11 rows, each a LayerNorm, two
skinny addmm and a ReLU.  LayerNorm and the FFN are row-wise, so chunking is the identity; here the
whole [B*T, dim] slab goes through one LayerNorm kernel and two tcgen05 GEMMs with bias / ReLU fused in the epilogue,
forward and backward hand-written (no autograd graph inside)."""
from __future__ import annotations

import torch

from . import ops
from .lsh_attention import _split_k


class _LNFeedForwardFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ln_w, ln_b, w1, b1, w2, b2, w1_bf16, w2_bf16, eps):
        shape = x.shape
        resid, resid_sub = ops.ResidualRequest.take(shape, x.device)      # reversible residual fused into the last GEMM's epilogue
        d = shape[-1]
        x2 = x.reshape(-1, d)
        if ln_w is not None:
            xn, mean, rstd = ops.layernorm_fwd(x2, ln_w, ln_b, eps)
        else:
            xn, mean, rstd = ops.cast_bf16_colsum(x2), None, None
        hid = ops.gemm(xn, w1_bf16, bias=b1, relu=True, out_dtype=torch.bfloat16)
        y = ops.gemm(hid, w2_bf16, bias=b2, resid=resid, resid_sub=resid_sub)
        ctx.has_ln = ln_w is not None
        ctx.params = (ln_w, ln_b, w1, b1, w2, b2)
        ctx.save_for_backward(x2, ln_w, mean, rstd, xn, hid, w1_bf16, w2_bf16)
        return y.view(shape)

    @staticmethod
    def backward(ctx, dy):
        x2, ln_w, mean, rstd, xn, hid, w1_bf16, w2_bf16 = ctx.saved_tensors
        rows, d = x2.shape
        f = hid.shape[1]
        dev = x2.device
        sk = _split_k(rows, (d // 128) * (f // 128))
        p_lnw, p_lnb, p_w1, p_b1, p_w2, p_b2 = ctx.params
        ctx.params = None

        def target(shape, param):
            """The parameter's own .grad inside the flat gradient buffer (the Function then returns None for it), else a zero-filled
            temporary that autograd adds (reformer_tts_b200.residual.grad_sink)."""
            sink = ops.grad_sink(param)
            return (sink, True) if sink is not None else (torch.zeros(shape, dtype=torch.float32, device=dev), False)

        g_b2, s_b2 = target((d,), p_b2)
        dyb = ops.cast_bf16_colsum(dy.reshape(rows, d), g_b2)
        g_w2, s_w2 = target((d, f), p_w2)
        ops.gemm(dyb, hid, a_mn_major=True, b_mn_major=True, out=g_w2, accumulate=True, split_k=sk)
        g_b1, s_b1 = target((f,), p_b1)
        # dh = (dy W2) * 1[h > 0]; W2 is [d, f] = [K, N] row-major -> MN-major B, no transposed copy
        dh = ops.gemm(dyb, w2_bf16, b_mn_major=True, gate=hid, colsum=g_b1, out_dtype=torch.bfloat16)
        g_w1, s_w1 = target((f, d), p_w1)
        ops.gemm(dh, xn, a_mn_major=True, b_mn_major=True, out=g_w1, accumulate=True, split_k=sk)
        dxn = ops.gemm(dh, w1_bf16, b_mn_major=True)
        g_lnw = g_lnb = None
        s_g = s_bt = False
        if ctx.has_ln:
            g_lnw, s_g = target((d,), p_lnw)
            g_lnb, s_bt = target((d,), p_lnb)
            dx = ops.layernorm_bwd(dxn, x2, ln_w, mean, rstd, g_lnw, g_lnb, accumulate_request=True)
        else:
            dx = dxn
        drop = lambda g, sunk: None if sunk else g
        return (dx.view(dy.shape), drop(g_lnw, s_g), drop(g_lnb, s_bt), drop(g_w1, s_w1), drop(g_b1, s_b1), drop(g_w2, s_w2), drop(g_b2, s_b2),
                None, None, None)


def ln_feed_forward(x, norm, lin1, lin2, cache1, cache2):
    if not x.is_cuda:
        raise RuntimeError("reformer_tts_b200 FeedForward runs on sm_100a CUDA only; there is no CPU path")
    ln_w = ln_b = None
    eps = 1e-5
    if norm is not None:
        ln_w, ln_b, eps = norm.weight, norm.bias, norm.eps
    return _LNFeedForwardFn.apply(x.float(), ln_w, ln_b, lin1.weight, lin1.bias, lin2.weight, lin2.bias,
                                  cache1.get(lin1.weight), cache2.get(lin2.weight), eps)


class _LayerNormFn(torch.autograd.Function):
    """LayerNorm on the row-wise kernels for wrapped layers that cannot take the norm into their own first kernel
    (the decoder's cross-attention): fp32 in, bf16 out (the consumer runs in bf16), fp32 statistics and gradients."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        y, mean, rstd = ops.layernorm_fwd(x2, weight, bias, eps)
        ctx.save_for_backward(x2, weight, mean, rstd)
        return y.view(shape)

    @staticmethod
    def backward(ctx, dy):
        x2, weight, mean, rstd = ctx.saved_tensors
        d = x2.shape[-1]
        g_w = torch.zeros(d, dtype=torch.float32, device=x2.device)
        g_b = torch.zeros(d, dtype=torch.float32, device=x2.device)
        dx = ops.layernorm_bwd(dy.reshape(-1, d).float(), x2, weight, mean, rstd, g_w, g_b)
        return dx.view(dy.shape), g_w, g_b, None


def layer_norm_bf16(x, norm):
    return _LayerNormFn.apply(x.float(), norm.weight, norm.bias, norm.eps)
