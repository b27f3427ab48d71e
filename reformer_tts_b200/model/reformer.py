"""Reformer encoder / decoder stacks with the interface of ref:reformer_tts/model/reformer.py, built on the
sm_100a LSH-attention and FeedForward layers.

Fusion points (results unchanged):
* ``WithNorm`` hands its LayerNorm to a wrapped layer that can run it inside its own first kernel
  (``forward_with_norm``) instead of launching ``nn.LayerNorm`` separately;
* ``Chunk`` calls a row-wise wrapped function once on the whole tensor: chunking a row-wise function along the
  row axis is the identity, and was only a memory device in the reference.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
from torch import nn

from ..cross_attention import cross_attention
from ..feed_forward import layer_norm_bf16
from ..lsh_attention import _WeightCache
from ..lsh_attention import HFLSHSelfAttention, LSHSelfAttention
from .modules import FeedForward
from .reversible import ReversibleBlock, ReversibleHalfResidual, ReversibleSequence, ReversibleSwap


class WithNorm(nn.Module):
    """``fn(norm(x))``  (ref:...reformer.py:25-33); state-dict keys ``norm.*`` / ``fn.*``."""

    def __init__(self, norm_class, dim, fn):
        super().__init__()
        self.norm = norm_class(dim)
        self.fn = fn

    @property
    def rowwise(self) -> bool:
        return bool(getattr(self.fn, "rowwise", False))

    def forward(self, x, **kwargs):
        if isinstance(self.norm, nn.LayerNorm) and hasattr(self.fn, "forward_with_norm"):
            return self.fn.forward_with_norm(x, self.norm, **kwargs)
        return self.fn(self.norm(x), **kwargs)


class Chunk(nn.Module):
    """``cat([fn(c) for c in x.chunk(chunks, dim)])``  (ref:...reformer.py:36-45)."""

    def __init__(self, chunks, fn, along_dim=-1):
        super().__init__()
        self.dim = along_dim
        self.chunks = chunks
        self.fn = fn

    def forward(self, x):
        if getattr(self.fn, "rowwise", False) and self.dim in (-2, x.dim() - 2):
            return self.fn(x)
        return torch.cat([self.fn(c) for c in x.chunk(self.chunks, dim=self.dim)], dim=self.dim)


class LSHSelfAttentionWrapper(nn.Module):
    """Selects the LSH implementation exactly like ref:...reformer.py:189-220; the layer lives at ``.layer`` so the
    state-dict keys are ``layer.{toqk,tov,to_out}.*`` or ``layer.{query_key,value}.weight``."""

    def __init__(self, dim: int, causal: bool, **kwargs):
        super().__init__()
        kwargs = dict(kwargs)
        implementation = kwargs.pop("implementation", None)
        if implementation not in {"huggingface_transformers", "reformer_pytorch"}:
            raise ValueError("attn kwargs need implementation in {'huggingface_transformers', 'reformer_pytorch'}")
        self.implementation = implementation
        if implementation == "reformer_pytorch":
            self.layer = LSHSelfAttention(dim, causal=causal, **kwargs)
        else:   # only these keys reach the HF config at ref:...reformer.py:204-212; everything else is ignored there too
            self.layer = HFLSHSelfAttention(dim, heads=kwargs["heads"], bucket_size=kwargs["bucket_size"],
                                            n_hashes=kwargs["n_hashes"], causal=causal, dropout=kwargs["dropout"])

    @property
    def takes_residual(self) -> bool:
        """The RP layer ends in the to_out GEMM (dropout and residual go into its epilogue); the HF layer has no output projection."""
        return self.implementation == "reformer_pytorch"

    def forward_with_norm(self, x, norm, input_mask: Optional[torch.Tensor] = None):
        if self.implementation == "reformer_pytorch":
            return self.layer(x, input_mask=input_mask, norm=norm)
        return self.layer(x, attention_mask=input_mask, norm=norm)

    def forward(self, x: torch.Tensor, input_mask: Optional[torch.Tensor] = None):
        return self.forward_with_norm(x, None, input_mask=input_mask)


class MultiheadAttentionWrapper(nn.Module):
    """Decoder-to-encoder attention (ref:...reformer.py:161-186): ``nn.MultiheadAttention`` parameters (state-dict keys
    ``layer.in_proj_weight`` ... ``layer.out_proj.bias``) with the argument order the reversible blocks need.  Outside the hot-path
    scope and kept on library kernels.  In training the attention weights are never read
    (ref:...reformer.py:183-184 appends them only in eval), so the training path skips materialising them: batch-first projections and
    ``scaled_dot_product_attention`` under bf16 autocast (same operand precision as the rest of the step); eval keeps the stock
    module call and returns the weights."""

    takes_residual = True      # kernel path only: the output projection can write x + f(.) / y - f(.) (ResidualRequest)

    def __init__(self, dim: int, attention_matrices: Optional[List[torch.Tensor]] = None, **kwargs):
        super().__init__()
        self.layer = nn.MultiheadAttention(dim, **kwargs)
        self.attention_matrices_ = attention_matrices
        self._caches = (_WeightCache(), _WeightCache(), _WeightCache())      # bf16 copies of Wq, Wkv, Wo

    def _fast_path_ok(self, query, extra) -> bool:
        layer = self.layer
        return (self.training and query.is_cuda and layer._qkv_same_embed_dim and layer.in_proj_bias is not None and layer.bias_k is None
                and not layer.add_zero_attn and set(extra) <= {"key_padding_mask"})

    def forward_with_norm(self, query, norm, **kwargs):
        """``WithNorm`` hands over its LayerNorm: in the training fast path it runs on the row-wise kernel (bf16 out)."""
        extra = {k: v for k, v in kwargs.items() if k not in ("key", "value")}
        if "key" in kwargs and self._fast_path_ok(query, extra) and query.shape[-1] % 128 == 0:
            memory = kwargs["key"]
            if self._kernel_path_ok(query, memory):
                return cross_attention(query, norm, memory, self.layer, self._caches, extra.get("key_padding_mask"))
            return self._forward_training(layer_norm_bf16(query, norm), memory, extra.get("key_padding_mask"))
        return self.forward(norm(query), **kwargs)

    def _kernel_path_ok(self, query, memory) -> bool:
        """Shapes the tcgen05 GEMMs take (rows in multiples of 128, head size 64); anything else stays on the library path."""
        b, t, d = query.shape
        return (memory.dim() == 3 and (b * t) % 128 == 0 and (b * memory.shape[1]) % 128 == 0 and d // self.layer.num_heads == 64
                and memory.shape[-1] == d)

    def forward(self, query, **kwargs):
        if "key" not in kwargs:
            raise AssertionError("forward expects keyword argument 'key'")
        extra = {k: v for k, v in kwargs.items() if k not in ("key", "value")}
        if self._fast_path_ok(query, extra):
            return self._forward_training(query, kwargs["key"], extra.get("key_padding_mask"))
        memory = kwargs["key"].transpose(0, 1)
        out, weights = self.layer(query.transpose(0, 1), memory, memory, **extra)
        if not self.training and self.attention_matrices_ is not None:
            self.attention_matrices_.append(weights)
        return out.transpose(0, 1)

    def _forward_training(self, query, memory, key_padding_mask):
        layer = self.layer
        b, t, d = query.shape
        h = layer.num_heads
        w, bias = layer.in_proj_weight, layer.in_proj_bias
        with torch.autocast("cuda", dtype=torch.bfloat16):
            q = nn.functional.linear(query, w[:d], bias[:d]).view(b, t, h, d // h).transpose(1, 2)
            kv = nn.functional.linear(memory, w[d:], bias[d:]).view(b, memory.shape[1], 2, h, d // h)
            k, v = kv[:, :, 0].transpose(1, 2), kv[:, :, 1].transpose(1, 2)
            mask = None if key_padding_mask is None else ~key_padding_mask[:, None, None, :]
            o = nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=layer.dropout)
            out = layer.out_proj(o.transpose(1, 2).reshape(b, t, d))
        return out.float()


def _normed_ff(dim, ff_chunks, ff_kwargs):
    ff = WithNorm(nn.LayerNorm, dim, FeedForward(dim, **ff_kwargs))
    return Chunk(ff_chunks, ff, along_dim=-2) if ff_chunks > 1 else ff


class ReformerEnc(nn.Module):
    """ref:...reformer.py:51-93: ``depth`` x ReversibleBlock(f = LN+LSH (non-causal), g = Chunk(LN+FF))."""

    def __init__(self, dim: int, depth: int, ff_chunks: int, attn_kwargs: Dict, ff_kwargs: Dict):
        super().__init__()
        self.dim, self.depth = dim, depth
        blocks = [ReversibleBlock(f=WithNorm(nn.LayerNorm, dim, LSHSelfAttentionWrapper(dim, causal=False, **attn_kwargs)),
                                  g=_normed_ff(dim, ff_chunks, ff_kwargs)) for _ in range(depth)]
        self.layers = ReversibleSequence(nn.ModuleList(blocks))

    def forward(self, x, input_mask=None, kwargs_list=None):
        if kwargs_list is None:
            kwargs_list = [dict() for _ in range(self.depth)]
        elif len(kwargs_list) != self.depth:
            raise AssertionError("list_kwargs should be the length of ReversibleSequence")
        for kwargs in kwargs_list:
            kwargs["f_args"] = {"input_mask": input_mask}
        y = self.layers(torch.cat([x, x], dim=-1), kwargs_list=kwargs_list)
        y1, y2 = y.chunk(2, dim=-1)
        return y1 + y2


class ReformerDec(nn.Module):
    """ref:...reformer.py:98-158: per layer HalfResidual(LN+LSH causal), Swap, HalfResidual(LN+cross-attention), Swap,
    HalfResidual(Chunk(LN+FF)), Swap; kwargs routed by stride-6 slices."""

    def __init__(self, dim: int, depth: int, ff_chunks: int, attn_kwargs: Dict, self_attn_kwargs: Dict, ff_kwargs: Dict):
        super().__init__()
        self.dim, self.depth = dim, depth
        self.attention_matrices_ = []
        blocks = []
        for _ in range(depth):
            blocks += [
                ReversibleHalfResidual(WithNorm(nn.LayerNorm, dim, LSHSelfAttentionWrapper(dim, causal=True, **self_attn_kwargs))),
                ReversibleSwap(),
                ReversibleHalfResidual(WithNorm(nn.LayerNorm, dim, MultiheadAttentionWrapper(dim, self.attention_matrices_, **attn_kwargs))),
                ReversibleSwap(),
                ReversibleHalfResidual(_normed_ff(dim, ff_chunks, ff_kwargs)),
                ReversibleSwap(),
            ]
        self.block_len = 6
        self.layers = ReversibleSequence(nn.ModuleList(blocks))

    def forward(self, x, keys, key_padding_mask=None, input_mask=None, kwargs_list=None):
        n = self.block_len * self.depth
        if kwargs_list is None:
            kwargs_list = [dict() for _ in range(n)]
        elif len(kwargs_list) != n:
            raise AssertionError("list_kwargs should be the length of ReversibleSequence")
        for kwargs in kwargs_list[2::6]:
            kwargs.update(key=keys, value=keys, key_padding_mask=key_padding_mask)
        for kwargs in kwargs_list[::6]:
            kwargs["input_mask"] = input_mask
        self.attention_matrices_.clear()
        y = self.layers(torch.cat([x, x], dim=-1), kwargs_list=kwargs_list)
        y1, y2 = y.chunk(2, dim=-1)
        return y1 + y2, self.attention_matrices_
