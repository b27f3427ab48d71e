"""Restatement of ``transformers`` ``LSHSelfAttention`` as the reference calls it (TEST INFRASTRUCTURE ONLY).

The reference builds the layer at ref:reformer_tts/model/reformer.py:201-213 and calls
``layer(x, attention_mask=input_mask).hidden_states`` at :218-220.  It pins transformers
2.11.0 (ref:requirements.txt:29); the only source present in the build image is 5.5.0, so
THIS FILE FOLLOWS 5.5.0 (``hf:`` line numbers below) and says so in every parity report
(version-drift row).  PINNED: ``tests/golden/make_golden.py`` ran the real
5.5.0 class here and committed its inputs/outputs; ``tests/test_oracle.py`` checks this
restatement against those vectors (buckets bit-equal, hidden states max-abs-diff 0 in fp32
on the generating machine, <=1e-5 elsewhere) and, when ``transformers`` is importable, live.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import lsh_core
from .lsh_core import LSHSpec


def auto_num_buckets(seq_len: int, chunk_len: int) -> int:
    """hf:781-785: 2 ** floor(log2(2 * (T // chunk))).  (The factorised-list branch at
    hf:788-793 needs > 2*max(sqrt(4096/chunk), chunk) buckets - never reached at T <= 16k
    with chunk >= 64 ... 128 buckets is the limit for chunk 64; see tests for the bound.)"""
    return 2 ** ((2 * (seq_len // chunk_len)).bit_length() - 1)


class LSHSelfAttentionHF(nn.Module):
    """Same parameters / state-dict keys as the real class: ``query_key.weight``, ``value.weight``."""

    def __init__(self, dim: int, heads: int, bucket_size: int, n_hashes: int, causal: bool, dropout: float = 0.):
        super().__init__()
        if dropout:
            raise NotImplementedError("attention-probability dropout is 0 in every reference config")
        self.dim, self.heads, self.chunk_len, self.n_hashes, self.causal = dim, heads, bucket_size, n_hashes, causal
        self.query_key = nn.Linear(dim, dim, bias=False)     # hf:428
        self.value = nn.Linear(dim, dim, bias=False)         # hf:429
        self.num_buckets = None                               # hf:531-533: set on first call, then cached
        self.last = None
        self.inject_buckets = None
        self.round_operands = False  # oracle/rounded.py: bf16 roundings at the CUDA path's operand boundaries

    def forward(self, x: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        b, t, _ = x.shape
        h, dh = self.heads, self.dim // self.heads
        if t <= self.chunk_len:
            raise NotImplementedError("hf:537-539 standard-attention fallback is outside the hot path")
        if self.round_operands:
            from . import rounded
            if self.num_buckets is None:
                self.num_buckets = auto_num_buckets(t, self.chunk_len)
            nb = self.num_buckets
            rot = torch.randn((h, dh, self.n_hashes, nb // 2), dtype=x.dtype, device=x.device)              # hf:717-719, same draw
            pad_bucket = attention_mask is not None and not bool(attention_mask.all())
            y, self.last = rounded.lsh_layer(x, self.query_key.weight, self.value.weight, None, None, self.inject_buckets, attention_mask,
                                             h, self.chunk_len, self.n_hashes, LSHSpec.huggingface(dh, self.causal), rot=rot,
                                             n_buckets=nb, pad_bucket=pad_bucket)
            self.last["rot"] = rot
            return y
        qk = self.query_key(x).view(b, t, h, dh).transpose(1, 2).reshape(b * h, t, dh)    # hf:511-524
        v = self.value(x).view(b, t, h, dh).transpose(1, 2).reshape(b * h, t, dh)
        if self.num_buckets is None:
            self.num_buckets = auto_num_buckets(t, self.chunk_len)
        nb = self.num_buckets
        mask = None if attention_mask is None else attention_mask.bool()[:, None, :].expand(b, h, t).reshape(b * h, t)
        # hf:717-719: one rotation per head, shared over the batch, global generator
        rot = torch.randn((h, dh, self.n_hashes, nb // 2), dtype=qk.dtype, device=qk.device)
        if self.inject_buckets is not None:
            buckets = self.inject_buckets.reshape(b * h, -1).long()
        else:
            rot_bh = rot[None].expand(b, -1, -1, -1, -1).reshape(b * h, dh, self.n_hashes, nb // 2)
            buckets = lsh_core.hash_buckets(qk, rot_bh, self.n_hashes, nb, pad_mask=mask)   # hf:720-758
        spec = LSHSpec.huggingface(dh, self.causal)
        res = lsh_core.lsh_attention(qk, v, buckets, self.chunk_len, self.n_hashes, spec, mask)
        res.update(qk=qk, v=v, rot=rot, buckets=buckets)
        self.last = res
        return res["out"].view(b, h, t, dh).transpose(1, 2).reshape(b, t, self.dim)       # hf:661


def build_real_hf_layer(dim: int, heads: int, bucket_size: int, n_hashes: int, causal: bool, dropout: float = 0.):
    """The real transformers class configured exactly as ref:reformer_tts/model/reformer.py:204-213
    (returns None when transformers is not importable)."""
    try:
        from transformers import ReformerConfig
        from transformers.models.reformer.modeling_reformer import LSHSelfAttention
    except Exception:       # pragma: no cover
        return None
    config = ReformerConfig(hidden_size=dim, is_decoder=causal, num_attention_heads=heads, num_hashes=n_hashes,
                            lsh_attention_probs_dropout_prob=dropout, lsh_attn_chunk_length=bucket_size,
                            attention_head_size=dim // heads)
    return LSHSelfAttention(config)
