"""Restatement of ``reformer_pytorch.LSHSelfAttention`` 0.19.1 (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED against the package itself: ``reformer-pytorch==0.19.1`` (pinned at
ref:requirements.txt:11, ref:setup.py:17) is not vendored under /root/reference and cannot be
installed here, and the reference holds no test or golden vector for it.
This file follows the published algorithm, rows R1-R11 of SURVEY.md 8(a); the call site it
stands in for is ref:reformer_tts/model/reformer.py:198-200,216-217 and its constructor
schema is ref:reformer_tts/model/config.py:10-27.  Pinned indirectly by tests/test_oracle.py
(KAT-1..4) and through the primitives shared with ``oracle.lsh_hf`` (which is pinned).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import lsh_core
from .lsh_core import LSHSpec


class LSHSelfAttentionRP(nn.Module):
    """Same parameters / state-dict keys as the package class: ``toqk.weight``, ``tov.weight``
    (no bias), ``to_out.{weight,bias}``.  Only the kwarg values the reference configs use are
    implemented; the rest raise."""

    def __init__(self, dim: int, heads: int = 8, bucket_size: int = 64, n_hashes: int = 8, causal: bool = False,
                 attn_chunks: int = 1, random_rotations_per_head: bool = False, attend_across_buckets: bool = True,
                 allow_duplicate_attention: bool = True, num_mem_kv: int = 0, one_value_head: bool = False,
                 use_full_attn: bool = False, full_attn_thres: Optional[int] = None, return_attn: bool = False,
                 post_attn_dropout: float = 0., dropout: float = 0., add_local_attn_hash: bool = False):
        super().__init__()
        if (random_rotations_per_head or not attend_across_buckets or not allow_duplicate_attention or num_mem_kv
                or one_value_head or use_full_attn or return_attn or dropout or add_local_attn_hash):
            raise NotImplementedError("oracle covers only the kwarg values used by the reference configs")
        assert dim % heads == 0
        self.dim, self.heads, self.bucket_size, self.n_hashes, self.causal = dim, heads, bucket_size, n_hashes, causal
        self.toqk = nn.Linear(dim, dim, bias=False)       # R1
        self.tov = nn.Linear(dim, dim, bias=False)
        self.to_out = nn.Linear(dim, dim)
        self.post_attn_dropout = nn.Dropout(post_attn_dropout)
        self.last = None    # stage outputs of the latest forward (tests read them)
        self.inject_buckets = None   # tests may force the bucket ids (SURVEY.md 8(c) end-to-end check)
        self.round_operands = False  # oracle/rounded.py: bf16 roundings at the CUDA path's operand boundaries

    def forward(self, x: torch.Tensor, input_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        b, t, _ = x.shape
        h, dh = self.heads, self.dim // self.heads
        assert t % (2 * self.bucket_size) == 0, "sequence must be a multiple of 2*bucket_size"
        if self.round_operands:
            from . import rounded
            n_buckets = t // self.bucket_size
            rot = torch.randn((1, dh, self.n_hashes, n_buckets // 2), dtype=x.dtype, device=x.device)      # R2, same draw
            y, self.last = rounded.lsh_layer(x, self.toqk.weight, self.tov.weight, self.to_out.weight, self.to_out.bias, self.inject_buckets,
                                             input_mask, h, self.bucket_size, self.n_hashes, LSHSpec.reformer_pytorch(dh, self.causal),
                                             rot=rot, n_buckets=n_buckets)
            self.last["rot"] = rot
            return self.post_attn_dropout(y)
        qk = self.toqk(x).view(b, t, h, dh).transpose(1, 2).reshape(b * h, t, dh)      # R1 merge_heads
        v = self.tov(x).view(b, t, h, dh).transpose(1, 2).reshape(b * h, t, dh)
        mask = None if input_mask is None else input_mask.bool()[:, None, :].expand(b, h, t).reshape(b * h, t)
        n_buckets = t // self.bucket_size
        # R2: ONE rotation tensor shared by every batch*head row, drawn from the global generator
        rot = torch.randn((1, dh, self.n_hashes, n_buckets // 2), dtype=qk.dtype, device=qk.device)
        if self.inject_buckets is not None:
            buckets = self.inject_buckets.reshape(b * h, -1).long()
        else:
            buckets = lsh_core.hash_buckets(qk, rot, self.n_hashes, n_buckets)
        spec = LSHSpec.reformer_pytorch(dh, self.causal)
        res = lsh_core.lsh_attention(qk, v, buckets, self.bucket_size, self.n_hashes, spec, mask)   # R3-R11
        res.update(qk=qk, v=v, rot=rot, buckets=buckets)
        self.last = res
        out = res["out"].view(b, h, t, dh).transpose(1, 2).reshape(b, t, self.dim)
        return self.post_attn_dropout(self.to_out(out))
