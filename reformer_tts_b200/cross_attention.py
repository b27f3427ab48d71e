"""Decoder-to-encoder attention block (ref:reformer_tts/model/reformer.py:161-186 under ``WithNorm``), training path.

SURVEY.md 8(f) rank 1: the first caller outside the LSH / FFN hot path that shares its reversible loop.  The block is one
autograd.Function with a hand-written backward, all of it on the library's own kernels: LayerNorm, every projection (forward,
dgrad, wgrad, bias gradients: row-wise kernels, tcgen05 GEMMs with fused bias / bf16 epilogues) and the dense softmax(QK^T)V core
over the <= 256 encoder positions (``rtts_xattn_fwd`` / ``rtts_xattn_bwd``, csrc/xattn.cu: key padding mask and the module's
attention-probability dropout inside the kernel, the mask regenerated from a device seed word drawn at the point of the call order
where ``nn.MultiheadAttention`` draws - so ``Deterministic``'s RNG replay and a CUDA-graph replay reproduce it in the recompute and
the backward; q / k / v / out stay token-major [B, T, D]: no head transposes).  Shapes the kernels do not take (T not a multiple of
128, S not a multiple of 64 or beyond 256) run the core on the vendor flash kernel (``scaled_dot_product_attention``) through a small
inner autograd graph.  Same parameters and state-dict keys as ``nn.MultiheadAttention`` (``in_proj_weight``, ``in_proj_bias``,
``out_proj.{weight,bias}``)."""
from __future__ import annotations

import torch
from torch.nn import functional as F

from . import ops
from .lsh_attention import _split_k


class _CrossAttentionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ln_w, ln_b, memory, w_in, b_in, w_out, b_out, wq_bf16, wkv_bf16, wo_bf16, keep_mask, seed, cfg):
        b, t, d = x.shape
        s = memory.shape[1]
        h, eps, p_drop = cfg["heads"], cfg["eps"], cfg["dropout"]
        dh = d // h
        x2 = x.reshape(b * t, d)
        if ln_w is not None:
            xn, mean, rstd = ops.layernorm_fwd(x2, ln_w, ln_b, eps)
        else:
            xn, mean, rstd = ops.cast_bf16_colsum(x2), None, None
        memb = ops.cast_bf16_colsum(memory.reshape(b * s, d))
        q = ops.gemm(xn, wq_bf16, bias=b_in[:d], out_dtype=torch.bfloat16)                 # [B*T, D]
        kv = ops.gemm(memb, wkv_bf16, bias=b_in[d:], out_dtype=torch.bfloat16)             # [B*S, 2D]
        own = ops.xattn_supported(t, s, dh) and s % 64 == 0
        need_grad = any(ctx.needs_input_grad)       # False in the reversible forward (no_grad)
        lse = keep_u8 = None
        if own:
            kv3 = kv.view(b, s, 2 * d)
            keep_u8 = None if keep_mask is None else keep_mask.reshape(b, s).to(torch.uint8)
            out3, lse = ops.xattn_fwd(q.view(b, t, d), kv3[..., :d], kv3[..., d:], keep_u8, h, dh ** -0.5, p_drop, seed)
            o2 = out3.view(b * t, d)
        else:
            kv5 = kv.view(b, s, 2, h, dh)
            with torch.set_grad_enabled(need_grad):     # vendor core: its backward is reached through this inner graph
                ql = q.view(b, t, h, dh).transpose(1, 2).detach().requires_grad_(need_grad)
                kl = kv5[:, :, 0].transpose(1, 2).detach().requires_grad_(need_grad)
                vl = kv5[:, :, 1].transpose(1, 2).detach().requires_grad_(need_grad)
                o = F.scaled_dot_product_attention(ql, kl, vl, attn_mask=keep_mask, dropout_p=p_drop)
            o2 = o.detach().transpose(1, 2).reshape(b * t, d)
            if not o2.is_contiguous():
                o2 = o2.contiguous()
        resid, resid_sub = ops.ResidualRequest.take(x.shape, x.device)     # reversible residual fused into the output projection
        y = ops.gemm(o2, wo_bf16, bias=b_out, resid=resid, resid_sub=resid_sub).view(b, t, d)
        ctx.inner = (ql, kl, vl, o) if need_grad and not own else None
        ctx.own = own
        ctx.p_drop = p_drop
        ctx.has_ln = ln_w is not None
        ctx.params = (ln_w, ln_b, w_in, w_out, b_out)
        ctx.dims = (b, t, s, d, h)
        ctx.save_for_backward(x2, ln_w, mean, rstd, xn, memb, o2, wq_bf16, wkv_bf16, wo_bf16, q if own else None, kv if own else None, lse, keep_u8,
                              seed if own else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x2, ln_w, mean, rstd, xn, memb, o2, wq_bf16, wkv_bf16, wo_bf16, q, kv, lse, keep_u8, seed = ctx.saved_tensors
        b, t, s, d, h = ctx.dims
        dh = d // h
        dev = x2.device
        rows_q, rows_kv = b * t, b * s
        tiles = (d // 128) ** 2
        p_lnw, p_lnb, p_win, p_wout, p_bout = ctx.params
        ctx.params = None

        def target(shape, param):
            """The parameter's own .grad inside the flat gradient buffer (then None is returned for it), else a zero-filled temporary
            that autograd adds (reformer_tts_b200.residual.grad_sink)."""
            sink = ops.grad_sink(param)
            return (sink, True) if sink is not None else (torch.zeros(shape, dtype=torch.float32, device=dev), False)

        # output projection
        g_bo, s_bo = target((d,), p_bout)
        dyb = ops.cast_bf16_colsum(dy.reshape(rows_q, d), g_bo)
        g_wo, s_wo = target((d, d), p_wout)
        ops.gemm(dyb, o2, a_mn_major=True, b_mn_major=True, out=g_wo, accumulate=True, split_k=_split_k(rows_q, tiles))
        do = ops.gemm(dyb, wo_bf16, b_mn_major=True, out_dtype=torch.bfloat16)            # [B*T, D]
        g_bin = torch.zeros(3 * d, dtype=torch.float32, device=dev)
        if ctx.own:
            # attention core: scores recomputed in-kernel from q, k and lse; dq complete per query tile, dk | dv summed over the tiles in fp32
            kv3 = kv.view(b, s, 2 * d)
            do3 = do.view(b, t, d)
            delta = ops.lsh_delta(do3, o2.view(b, t, d), h)
            dq3, dkv32 = ops.xattn_bwd(q.view(b, t, d), kv3[..., :d], kv3[..., d:], keep_u8, h, dh ** -0.5, ctx.p_drop, seed, do3, lse, delta)
            dq = dq3.view(rows_q, d)
            dkv = ops.cast_bf16_colsum(dkv32.view(rows_kv, 2 * d), g_bin[d:])
        else:
            # attention core (vendor kernel, saved inner graph)
            ql, kl, vl, o = ctx.inner
            ctx.inner = None
            dq4, dk4, dv4 = torch.autograd.grad(o, (ql, kl, vl), do.view(b, t, h, dh).transpose(1, 2))
            dq = dq4.transpose(1, 2).reshape(rows_q, d)
            if not dq.is_contiguous():
                dq = dq.contiguous()
            dkv = torch.stack((dk4.transpose(1, 2), dv4.transpose(1, 2)), dim=2).reshape(rows_kv, 2 * d)
            ops.colsum_bf16(dkv, g_bin[d:])
        ops.colsum_bf16(dq, g_bin[:d])
        # input projections: weights [3D, D] = [Wq; Wk; Wv]
        g_win, s_win = target((3 * d, d), p_win)
        ops.gemm(dq, xn, a_mn_major=True, b_mn_major=True, out=g_win[:d], accumulate=True, split_k=_split_k(rows_q, tiles))
        ops.gemm(dkv, memb, a_mn_major=True, b_mn_major=True, out=g_win[d:], accumulate=True, split_k=_split_k(rows_kv, 2 * tiles))
        dxn = ops.gemm(dq, wq_bf16, b_mn_major=True)                                        # fp32 [B*T, D]
        dmem = ops.gemm(dkv, wkv_bf16, b_mn_major=True).view(b, s, d)                       # fp32 [B, S, D]
        g_lnw = g_lnb = None
        s_g = s_bt = False
        if ctx.has_ln:
            g_lnw, s_g = target((d,), p_lnw)
            g_lnb, s_bt = target((d,), p_lnb)
            dx = ops.layernorm_bwd(dxn, x2, ln_w, mean, rstd, g_lnw, g_lnb, accumulate_request=True)
        else:
            dx = dxn
        drop = lambda g, sunk: None if sunk else g
        return (dx.view(b, t, d), drop(g_lnw, s_g), drop(g_lnb, s_bt), dmem, drop(g_win, s_win), g_bin, drop(g_wo, s_wo), drop(g_bo, s_bo),
                None, None, None, None, None, None)


def cross_attention(x, norm, memory, layer, caches, key_padding_mask):
    """x fp32 [B,T,D] (un-normalised query), memory fp32 [B,S,D]; ``layer`` = the wrapped nn.MultiheadAttention."""
    if not x.is_cuda:
        raise RuntimeError("reformer_tts_b200 cross-attention training path runs on sm_100a CUDA only")
    d = x.shape[-1]
    w_in, b_in = layer.in_proj_weight, layer.in_proj_bias
    ln_w = ln_b = None
    eps = 1e-5
    if norm is not None:
        ln_w, ln_b, eps = norm.weight, norm.bias, norm.eps
    keep = None if key_padding_mask is None else ~key_padding_mask[:, None, None, :]
    cfg = dict(heads=layer.num_heads, eps=eps, dropout=layer.dropout if layer.training else 0.0)
    # the seed word of the in-kernel dropout mask: drawn here, where nn.MultiheadAttention would draw its mask, from the current CUDA
    # generator (Deterministic replays the generator state in the recompute, so the same word - hence the same mask - comes out again)
    seed = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, device=x.device) if cfg["dropout"] > 0.0 else None
    return _CrossAttentionFn.apply(x.float(), ln_w, ln_b, memory.float(), w_in, b_in, layer.out_proj.weight, layer.out_proj.bias,
                                   caches[0].get(w_in[:d]), caches[1].get(w_in[d:]), caches[2].get(layer.out_proj.weight), keep, seed, cfg)
