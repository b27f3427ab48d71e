"""Shared primitives of the two LSH-attention restatements (TEST INFRASTRUCTURE ONLY).

Every function is plain differentiable PyTorch on CPU (fp32 or fp64), written from
the algorithm description in SURVEY.md 3.2 / 8(a); citations give the place in the
reference (or in the third-party file the reference calls into) that the step follows:

  ref: = /root/reference/            (kowaalczyk/reformer-tts)
  hf:  = transformers/models/reformer/modeling_reformer.py  (5.5.0, in the build image)
  rp:  = reformer-pytorch 0.19.1 ``reformer_pytorch/reformer_pytorch.py`` (not on the box;
         step numbers R1..R11 are the rows of SURVEY.md 8(a))

The same ``LSHSpec`` constants are what the CUDA kernels take through the C ABI
(include/rtts_b200.h ``rtts_lsh_spec``), so kernel and oracle are parametrised alike.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

KEYNORM_L2 = 0        # rp R5: F.normalize(x, p=2, dim=-1), eps 1e-12
KEYNORM_RMS = 1       # hf:1042-1056: x * rsqrt(mean(x^2) + 1e-6) / sqrt(dh)
MASK_QUERY_AND_KEY = 0  # rp R8: masked when query OR key is padding
MASK_KEY_ONLY = 1       # hf:914-922: masked when the key is padding


@dataclass(frozen=True)
class LSHSpec:
    """Constants that differ between the two third-party implementations."""

    score_scale: float      # multiplies q.k        (rp R7: dh**-0.5; hf: 1.0, folded into key norm)
    key_norm: int           # KEYNORM_*
    mask_value: float       # fill for padding / causal   (rp R8: -finfo.max; hf:431 -1e9)
    self_value: float       # fill for q_pos == k_pos     (rp R8: -5e4;       hf:430 -1e5)
    mask_mode: int          # MASK_*
    causal: bool

    @staticmethod
    def reformer_pytorch(dh: int, causal: bool) -> "LSHSpec":
        return LSHSpec(dh ** -0.5, KEYNORM_L2, -torch.finfo(torch.float32).max, -5e4,
                       MASK_QUERY_AND_KEY, causal)

    @staticmethod
    def huggingface(dh: int, causal: bool) -> "LSHSpec":
        return LSHSpec(1.0, KEYNORM_RMS, -1e9, -1e5, MASK_KEY_ONLY, causal)


def hash_buckets(qk: torch.Tensor, rot: torch.Tensor, n_rounds: int, n_buckets: int,
                 pad_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Random-rotation LSH (rp R2 / hf:717-758).

    qk  [N, T, dh]; rot [N or 1, dh, R, n_buckets/2] (rp shares one rotation over all N
    = batch*heads rows, hf passes one per head, expanded by the caller).
    Returns int64 buckets [N, R*T] with round offsets already added.
    ``pad_mask`` [N, T] (True = real token) reproduces hf:740-747: padded tokens go to an
    extra bucket ``n_buckets`` and the round offset becomes ``n_buckets + 1``; the extra
    bucket is only introduced when at least one token in the whole mask is padding.
    """
    n, t, dh = qk.shape
    assert n_buckets % 2 == 0 and rot.shape[1:] == (dh, n_rounds, n_buckets // 2)
    rot = rot.expand(n, -1, -1, -1)
    proj = torch.einsum("ntd,ndri->nrti", qk.detach(), rot)
    ids = torch.argmax(torch.cat([proj, -proj], dim=-1), dim=-1)          # [N, R, T]
    stride = n_buckets
    if pad_mask is not None and not bool(pad_mask.all()):
        stride = n_buckets + 1
        ids = torch.where(pad_mask[:, None, :], ids, torch.full_like(ids, n_buckets))
    ids = ids + stride * torch.arange(n_rounds, device=qk.device).view(1, -1, 1)
    return ids.reshape(n, n_rounds * t)


def sort_buckets(buckets: torch.Tensor, seq_len: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Stable sort of (bucket, position) (rp R3 / hf:150-156,762-779).

    Returns (sticker, undo) int64 [N, R*T]: ``sticker[i]`` is the flat index (round*T + pos)
    sitting at sorted slot ``i``; ``undo[sticker[i]] == i``.  Keys are unique, so any correct
    sort gives the same permutation as ``argsort(buckets, stable=True)``.
    """
    n, total = buckets.shape
    ticker = torch.arange(total, device=buckets.device).expand(n, -1)
    keys = seq_len * buckets + ticker % seq_len
    sticker = torch.argsort(keys, dim=-1)
    undo = torch.empty_like(sticker)
    undo.scatter_(-1, sticker, ticker)
    return sticker, undo


def _with_previous_chunk(x: torch.Tensor) -> torch.Tensor:
    """Look-one-back (rp R6 / hf:352-373): chunk c sees [c, c-1]; chunk 0 wraps to the last."""
    return torch.cat([x, torch.roll(x, shifts=1, dims=1)], dim=2)


def normalise_keys(x: torch.Tensor, spec: LSHSpec) -> torch.Tensor:
    if spec.key_norm == KEYNORM_L2:
        return x / x.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    var = (x * x).mean(dim=-1, keepdim=True)
    return x * torch.rsqrt(var + 1e-6) / math.sqrt(x.shape[-1])


def chunk_attention(qk: torch.Tensor, v: torch.Tensor, sticker: torch.Tensor, bucket_size: int,
                    n_rounds: int, spec: LSHSpec, mask: Optional[torch.Tensor] = None
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Sorted, chunked shared-QK attention with look-one-back (rp R4-R9 / hf:563-599,801-906).

    qk, v [N, T, dh]; sticker [N, R*T]; mask [N, T] bool (True = real token) or None.
    Returns the per-slot outputs in SORTED order: so [N, R*T, dh], slse [N, R*T].
    """
    n, t, dh = qk.shape
    st = sticker % t
    chunks = n_rounds * t // bucket_size
    idx = st.unsqueeze(-1).expand(-1, -1, dh)
    bq = qk.gather(1, idx).reshape(n, chunks, bucket_size, dh)
    bv = v.gather(1, idx).reshape(n, chunks, bucket_size, dh)
    bk = normalise_keys(bq, spec)
    q_pos = st.reshape(n, chunks, bucket_size)
    bk, bv, k_pos = _with_previous_chunk(bk), _with_previous_chunk(bv), _with_previous_chunk(q_pos)

    dots = torch.einsum("ncid,ncjd->ncij", bq, bk) * spec.score_scale
    fill = torch.tensor(spec.mask_value, dtype=dots.dtype)
    if mask is not None:
        mq = mask.gather(1, st).reshape(n, chunks, bucket_size)
        mk = _with_previous_chunk(mq)
        keep = mk[:, :, None, :] if spec.mask_mode == MASK_KEY_ONLY else mq[:, :, :, None] & mk[:, :, None, :]
        dots = torch.where(keep, dots, fill)
    if spec.causal:
        dots = torch.where(q_pos[:, :, :, None] < k_pos[:, :, None, :], fill, dots)
    dots = torch.where(q_pos[:, :, :, None] == k_pos[:, :, None, :],
                       torch.tensor(spec.self_value, dtype=dots.dtype), dots)

    lse = torch.logsumexp(dots, dim=-1, keepdim=True)
    probs = torch.exp(dots - lse)
    so = torch.einsum("ncij,ncjd->ncid", probs, bv)
    return so.reshape(n, -1, dh), lse.reshape(n, -1)


def unsort_and_merge(so: torch.Tensor, slse: torch.Tensor, undo: torch.Tensor, n_rounds: int
                     ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Undo the sort and combine hash rounds (rp R10-R11 / hf:1067-1096,626-645).

    Returns out [N, T, dh], per-round o [N, R, T, dh], per-round lse [N, R, T].
    """
    n, total, dh = so.shape
    t = total // n_rounds
    o = so.gather(1, undo.unsqueeze(-1).expand(-1, -1, dh)).reshape(n, n_rounds, t, dh)
    lse = slse.gather(1, undo).reshape(n, n_rounds, t)
    w = torch.exp(lse - torch.logsumexp(lse, dim=1, keepdim=True))
    return (o * w.unsqueeze(-1)).sum(dim=1), o, lse


def lsh_attention(qk: torch.Tensor, v: torch.Tensor, buckets: torch.Tensor, bucket_size: int,
                  n_rounds: int, spec: LSHSpec, mask: Optional[torch.Tensor] = None):
    """hash output -> attention output.  Returns dict with every observable stage."""
    t = qk.shape[1]
    sticker, undo = sort_buckets(buckets, t)
    so, slse = chunk_attention(qk, v, sticker, bucket_size, n_rounds, spec, mask)
    out, o_rounds, lse_rounds = unsort_and_merge(so, slse, undo, n_rounds)
    return {"out": out, "sticker": sticker, "undo": undo, "o_rounds": o_rounds,
            "lse_rounds": lse_rounds, "lse": torch.logsumexp(lse_rounds, dim=1)}


def dense_shared_qk_attention(qk: torch.Tensor, v: torch.Tensor, spec: LSHSpec,
                              mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Full (un-hashed) shared-QK attention with the same masking rules; used by the
    known-answer test KAT-1 (n_buckets == 2, one round => LSH attention == dense attention)."""
    n, t, dh = qk.shape
    dots = torch.einsum("nid,njd->nij", qk, normalise_keys(qk, spec)) * spec.score_scale
    pos = torch.arange(t)
    fill = torch.tensor(spec.mask_value, dtype=dots.dtype)
    if mask is not None:
        keep = mask[:, None, :] if spec.mask_mode == MASK_KEY_ONLY else mask[:, :, None] & mask[:, None, :]
        dots = torch.where(keep, dots, fill)
    if spec.causal:
        dots = torch.where(pos[:, None] < pos[None, :], fill, dots)
    dots = torch.where(pos[:, None] == pos[None, :], torch.tensor(spec.self_value, dtype=dots.dtype), dots)
    return torch.softmax(dots, dim=-1) @ v
