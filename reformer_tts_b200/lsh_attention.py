"""LSH self-attention layers with the constructor / forward API of the two third-party classes the
reference can select (ref:reformer_tts/model/reformer.py:198-213), running on the sm_100a kernels.

* ``LSHSelfAttention``    - ``reformer_pytorch.LSHSelfAttention`` 0.19.1 API (rp R1-R11):
  parameters ``toqk.weight``, ``tov.weight`` (no bias), ``to_out.{weight,bias}``.
* ``HFLSHSelfAttention``  - ``transformers`` ``LSHSelfAttention`` as configured at
  ref:reformer_tts/model/reformer.py:204-213: parameters ``query_key.weight``, ``value.weight``.

Both are thin shells around one autograd.Function whose forward AND backward are hand-written kernel
sequences (no autograd graph inside the layer, so the reversible recompute re-runs exactly the fused
forward).  Rotations are drawn with ``torch.randn`` on the compute device at the same point of the call
order and with the same shape as the third-party code, so ``Deterministic`` RNG replay reproduces them.
There is no non-CUDA path: construction is cheap, but ``forward`` on a CPU tensor raises.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn
import weakref

from . import ops
from .ops import LSHSpec


class _WeightCache:
    """bf16 copies of fp32 master weights in a persistent buffer, rebuilt only when a weight's version counter moves.
    ``enabled = False`` forces the cast on every call.  ``refresh_all()`` re-casts every known weight of the process with ONE
    multi-tensor copy (the per-weight cat + cast + copy was ~130 small launches per training step); a training step calls it
    right after bumping ``epoch`` (inside the captured graph too: a replay updates the weights without running Python)."""
    enabled = True
    epoch = 0       # bumped at the start of every training step: forces one re-cast per step even when version counters cannot be observed
    _registry = weakref.WeakSet()

    def __init__(self):
        self._key = None
        self._val = None
        self._parts = None
        self._weights = None
        _WeightCache._registry.add(self)

    def _make_key(self, weights):
        return (_WeightCache.epoch,) + tuple((w.data_ptr(), w._version) for w in weights)

    def _ensure_buffer(self, weights):
        rows = sum(w.shape[0] for w in weights)
        if (self._val is None or self._val.shape[0] != rows or self._val.shape[1:] != weights[0].shape[1:] or self._val.device != weights[0].device
                or self._weights is None or len(self._weights) != len(weights) or any(a is not b for a, b in zip(self._weights, weights))):
            self._val = torch.empty((rows,) + tuple(weights[0].shape[1:]), dtype=torch.bfloat16, device=weights[0].device)
            self._parts, r = [], 0
            for w in weights:
                self._parts.append(self._val[r:r + w.shape[0]])
                r += w.shape[0]
            self._weights = tuple(weights)

    def get(self, *weights: torch.Tensor) -> torch.Tensor:
        key = self._make_key(weights)
        if key != self._key or not _WeightCache.enabled:
            with torch.no_grad():
                self._ensure_buffer(weights)
                torch._foreach_copy_(self._parts, [w.detach() for w in weights])
            self._key = key
        return self._val

    @classmethod
    def refresh_all(cls):
        """Re-cast every weight that has been requested before (one multi-tensor kernel per dtype / device group)."""
        dst, src = [], []
        for c in list(cls._registry):
            if c._weights is None or c._val is None:
                continue
            dst += c._parts
            src += [w.detach() for w in c._weights]
            c._key = c._make_key(c._weights)
        if dst:
            with torch.no_grad():
                torch._foreach_copy_(dst, src)


class _LSHAttentionFn(torch.autograd.Function):
    """x fp32 [B,T,D] -> y fp32 [B,T,D]:  (LayerNorm) -> QK|V projection -> hash -> sort -> chunked attention
    -> round merge -> (output projection).  Saves only x, LN statistics, the bf16 projections, the sort and
    the merged output; scores are recomputed in the backward kernel."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, w_qk, w_v, w_out, b_out, wqkv_bf16, wout_bf16, rot, mask_u8, cfg):
        b, t, d = x.shape
        h, r, bucket, spec, nb, pad_bucket = cfg["heads"], cfg["n_hashes"], cfg["bucket_size"], cfg["spec"], cfg["n_buckets"], cfg["pad_bucket"]
        x2 = x.reshape(b * t, d)
        if ln_w is not None:
            xn, mean, rstd = ops.layernorm_fwd(x2, ln_w, ln_b, cfg["eps"])
        else:
            xn, mean, rstd = ops.cast_bf16_colsum(x2), None, None
        qkv = ops.gemm(xn, wqkv_bf16, out_dtype=torch.bfloat16).view(b, t, 2 * d)
        qk, v = qkv[..., :d], qkv[..., d:]
        buckets, sumsq = ops.lsh_hash(qk, rot, h, r, nb, mask_u8 if pad_bucket else None, pad_bucket, return_sumsq=True)
        sticker, undo = ops.lsh_sort(buckets, t, r, nb + 1 if pad_bucket else nb)
        o_rounds, lse_rounds = ops.lsh_attn_fwd(qk, v, sticker, mask_u8, spec, h, r, bucket, sumsq=sumsq)
        out, lse = ops.lsh_merge_fwd(o_rounds, lse_rounds)
        keep, keep_scale = cfg.get("keep_mask"), cfg.get("keep_scale", 1.0)
        if wout_bf16 is not None:
            # output projection with post_attn_dropout (inverted-dropout keep mask) and the reversible residual in its epilogue
            resid, resid_sub = ops.ResidualRequest.take(x.shape, x.device)
            y = ops.gemm(out.view(b * t, d), wout_bf16, bias=b_out, resid=resid, resid_sub=resid_sub,
                         keep_mask=None if keep is None else keep.view(b * t, d), keep_scale=keep_scale).view(b, t, d)
        else:
            y = out.float()
        ctx.cfg = cfg
        ctx.has_ln, ctx.has_out = ln_w is not None, wout_bf16 is not None
        ctx.params = (ln_w, ln_b, w_qk, w_v, w_out, b_out)
        ctx.save_for_backward(x, ln_w, mean, rstd, xn, qkv, sticker, undo, out, lse, mask_u8, wqkv_bf16, wout_bf16, sumsq)
        cfg["_last_buckets"] = buckets
        return y

    @staticmethod
    def backward(ctx, dy):
        x, ln_w, mean, rstd, xn, qkv, sticker, undo, out, lse, mask_u8, wqkv_bf16, wout_bf16, sumsq = ctx.saved_tensors
        cfg = ctx.cfg
        b, t, d = x.shape
        h, r, bucket, spec = cfg["heads"], cfg["n_hashes"], cfg["bucket_size"], cfg["spec"]
        dev = x.device
        dy2 = dy.reshape(b * t, d)
        p_lnw, p_lnb, p_wqk, p_wv, p_wout, p_bout = ctx.params
        ctx.params = None

        def target(shape, *params):
            """Where a parameter gradient is accumulated: the parameters' own .grad inside the flat gradient buffer (then the
            Function returns None for them) or a zero-filled temporary that autograd adds."""
            sink = ops.grad_sink(*params)
            return (sink, True) if sink is not None else (torch.zeros(shape, dtype=torch.float32, device=dev), False)

        g_wout = g_bout = None
        if ctx.has_out:
            g_bout, sunk_b = target((d,), p_bout)
            keep = cfg.get("keep_mask")       # the forward's dropout mask: d(dropout(z)) = keep * scale * dy
            dyb = ops.cast_bf16_colsum(dy2, g_bout, keep_mask=None if keep is None else keep.view(b * t, d), keep_scale=cfg.get("keep_scale", 1.0))
            g_wout, sunk_w = target((d, d), p_wout)
            ops.gemm(dyb, out.view(b * t, d), a_mn_major=True, b_mn_major=True, out=g_wout, accumulate=True, split_k=_split_k(b * t, (d // 128) ** 2))
            dout = ops.gemm(dyb, wout_bf16, b_mn_major=True, out_dtype=torch.bfloat16).view(b, t, d)
            g_bout, g_wout = (None if sunk_b else g_bout), (None if sunk_w else g_wout)
        else:
            dout = ops.cast_bf16_colsum(dy2).view(b, t, d)
        delta = ops.lsh_delta(dout, out, h)
        qk, v = qkv[..., :d], qkv[..., d:]
        dqkv = torch.empty((b, t, 2 * d), dtype=torch.bfloat16, device=dev)
        ops.lsh_attn_bwd(qk, v, sticker, undo, mask_u8, spec, dout, lse, delta, h, r, bucket, out_dqk=dqkv[..., :d], out_dv=dqkv[..., d:], sumsq=sumsq)
        dqkv2 = dqkv.view(b * t, 2 * d)
        g_wqkv, sunk_qkv = target((2 * d, d), p_wqk, p_wv)
        ops.gemm(dqkv2, xn, a_mn_major=True, b_mn_major=True, out=g_wqkv, accumulate=True, split_k=_split_k(b * t, 2 * (d // 128) ** 2))
        g_wqk, g_wv = (None, None) if sunk_qkv else (g_wqkv[:d], g_wqkv[d:])
        dxn = ops.gemm(dqkv2, wqkv_bf16, b_mn_major=True)
        g_lnw = g_lnb = None
        if ctx.has_ln:
            g_lnw, sunk_g = target((d,), p_lnw)
            g_lnb, sunk_bt = target((d,), p_lnb)
            dx = ops.layernorm_bwd(dxn, x.reshape(b * t, d), ln_w, mean, rstd, g_lnw, g_lnb, accumulate_request=True).view(b, t, d)
            g_lnw, g_lnb = (None if sunk_g else g_lnw), (None if sunk_bt else g_lnb)
        else:
            dx = dxn.view(b, t, d)
        return dx, g_lnw, g_lnb, g_wqk, g_wv, g_wout, g_bout, None, None, None, None, None


def _split_k(tokens: int, out_tiles: int = 64) -> int:
    """Split factor for the weight-gradient GEMMs (K = tokens, ``out_tiles`` 128x128 output tiles): about two CTAs per SM, and
    as few splits as that allows - every split adds an fp32 atomic pass over the output."""
    kb = tokens // 64
    best = 1
    for s in (1, 2, 4, 5, 8, 10, 16, 20, 32, 40):
        if kb % s == 0 and kb // s >= 4:
            best = s
            if s * out_tiles >= 256:
                break
    return best


class _LSHBase(nn.Module):
    _last = (None, None, 1)      # (bucket ids, padding mask if the extra bucket was used, rounds) of the latest forward

    def _run(self, x, norm, w_qk, w_v, w_out, b_out, rot, mask, cfg):
        if not x.is_cuda:
            raise RuntimeError("reformer_tts_b200 LSH attention runs on sm_100a CUDA only; there is no CPU path")
        wqkv = self._wqkv_cache.get(w_qk, w_v)
        wout = self._wout_cache.get(w_out) if w_out is not None else None
        mask_u8 = None if mask is None else mask.to(device=x.device, dtype=torch.uint8).contiguous()
        ln_w = ln_b = None
        if norm is not None:
            ln_w, ln_b = norm.weight, norm.bias
            cfg = dict(cfg, eps=norm.eps)
        y = _LSHAttentionFn.apply(x.float(), ln_w, ln_b, w_qk, w_v, w_out, b_out, wqkv, wout, rot, mask_u8, cfg)
        self._last = (cfg.pop("_last_buckets", None), mask_u8 if cfg["pad_bucket"] else None, cfg["n_hashes"])
        return y

    @property
    def last_buckets(self):
        """int32 [B,H,R*T] bucket ids of the latest forward in the numbering of the third-party class (parity tests read it).
        The HF layer always hashes with the extra padding bucket when a mask is supplied (round stride nb + 1); transformers does
        so only if the mask actually masks something (hf:740-747) - the sort order is the same either way, the ids are converted
        here, on demand, off the hot path."""
        buckets, pad_mask, rounds = self._last
        if buckets is not None and pad_mask is not None and bool(pad_mask.all()):
            t = buckets.shape[-1] // rounds
            buckets = buckets - torch.arange(rounds, device=buckets.device, dtype=buckets.dtype).repeat_interleave(t)
        return buckets


class LSHSelfAttention(_LSHBase):
    """Drop-in for ``reformer_pytorch.LSHSelfAttention`` (0.19.1) at the kwarg values the reference configs use
    (SURVEY.md 8(b)); any other value raises NotImplementedError at construction - no silent fallback."""

    def __init__(self, dim, heads=8, bucket_size=64, n_hashes=8, causal=False, attn_chunks=1, random_rotations_per_head=False,
                 attend_across_buckets=True, allow_duplicate_attention=True, num_mem_kv=0, one_value_head=False,
                 use_full_attn=False, full_attn_thres=None, return_attn=False, post_attn_dropout=0., dropout=0.,
                 add_local_attn_hash=False, **unused):
        super().__init__()
        unsupported = dict(random_rotations_per_head=random_rotations_per_head, attend_across_buckets=not attend_across_buckets,
                           allow_duplicate_attention=not allow_duplicate_attention, num_mem_kv=num_mem_kv,
                           one_value_head=one_value_head, use_full_attn=use_full_attn, return_attn=return_attn,
                           dropout=dropout, add_local_attn_hash=add_local_attn_hash, attn_chunks=attn_chunks != 1)
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError(f"LSHSelfAttention: unsupported option(s) {bad}; only the reference configs' values are built")
        if dim % heads or dim // heads != 64:
            raise NotImplementedError("LSHSelfAttention: head size must be 64 (dim // heads)")
        if bucket_size not in (64, 128):
            raise NotImplementedError("LSHSelfAttention: bucket_size must be 64 or 128")
        self.dim, self.heads, self.bucket_size, self.n_hashes, self.causal = dim, heads, bucket_size, n_hashes, causal
        self.full_attn_thres = bucket_size if full_attn_thres is None else full_attn_thres
        self.toqk = nn.Linear(dim, dim, bias=False)
        self.tov = nn.Linear(dim, dim, bias=False)
        self.to_out = nn.Linear(dim, dim)
        self.post_attn_dropout = nn.Dropout(post_attn_dropout)
        self._wqkv_cache, self._wout_cache = _WeightCache(), _WeightCache()
        self.rot_override = None

    def forward(self, x, input_mask=None, norm: Optional[nn.LayerNorm] = None, **kwargs):
        b, t, d = x.shape
        if t % (2 * self.bucket_size) != 0:
            raise ValueError(f"sequence length {t} must be divisible by 2*bucket_size = {2 * self.bucket_size}")
        if t <= self.full_attn_thres:
            raise NotImplementedError("full-attention fallback (T <= full_attn_thres) is outside the hot path")
        n_buckets = t // self.bucket_size
        # rp R2: one rotation set shared by all batch*head rows, from the global generator of the compute device
        rot = torch.randn((1, d // self.heads, self.n_hashes, n_buckets // 2), dtype=torch.float32, device=x.device)
        if self.rot_override is not None:       # parity tests inject the oracle's rotations (CPU and CUDA generators differ)
            rot = self.rot_override.to(device=x.device, dtype=torch.float32)
        cfg = dict(heads=self.heads, n_hashes=self.n_hashes, bucket_size=self.bucket_size, n_buckets=n_buckets, pad_bucket=False,
                   spec=LSHSpec.reformer_pytorch(d // self.heads, self.causal), eps=1e-5)
        # post_attn_dropout (nn.Dropout after to_out): the keep mask is drawn here, at the point of the call order where nn.Dropout
        # would draw it (so Deterministic's RNG replay reproduces it in the recompute), and applied in the to_out GEMM's epilogue
        p_drop = self.post_attn_dropout.p
        if self.training and p_drop > 0.:
            cfg["keep_mask"] = torch.empty((b, t, d), dtype=torch.uint8, device=x.device).bernoulli_(1. - p_drop)
            cfg["keep_scale"] = 1. / (1. - p_drop)
        return self._run(x, norm, self.toqk.weight, self.tov.weight, self.to_out.weight, self.to_out.bias, rot, input_mask, cfg)


class HFLSHSelfAttention(_LSHBase):
    """Drop-in for ``transformers`` ``LSHSelfAttention`` as the reference configures it; ``forward`` returns the
    hidden states tensor directly (the wrapper takes ``.hidden_states``)."""

    def __init__(self, dim, heads, bucket_size, n_hashes, causal, dropout=0.):
        super().__init__()
        if dropout:
            raise NotImplementedError("HFLSHSelfAttention: attention-probability dropout is not built (0 in every reference config)")
        if dim % heads or dim // heads != 64:
            raise NotImplementedError("HFLSHSelfAttention: head size must be 64")
        if bucket_size not in (64, 128):
            raise NotImplementedError("HFLSHSelfAttention: lsh_attn_chunk_length must be 64 or 128")
        self.dim, self.heads, self.bucket_size, self.n_hashes, self.causal = dim, heads, bucket_size, n_hashes, causal
        self.query_key = nn.Linear(dim, dim, bias=False)
        self.value = nn.Linear(dim, dim, bias=False)
        self.num_buckets = None      # hf:531-533 set lazily on the first call, then kept
        self._wqkv_cache, self._wout_cache = _WeightCache(), _WeightCache()
        self.rot_override = None
        self.max_position_embeddings = 4096      # ReformerConfig default; the reference does not set it (ref:...reformer.py:204-212)

    def forward(self, x, attention_mask=None, norm: Optional[nn.LayerNorm] = None, **kwargs):
        b, t, d = x.shape
        if t <= self.bucket_size:
            raise NotImplementedError("hf:537-539 standard-attention fallback is outside the hot path")
        if t % (2 * self.bucket_size) != 0:
            raise ValueError(f"sequence length {t} must be divisible by 2*chunk_length = {2 * self.bucket_size}")
        if self.num_buckets is None:
            self.num_buckets = 2 ** ((2 * (t // self.bucket_size)).bit_length() - 1)      # hf:781-785
        nb = self.num_buckets
        # hf:788-793: above this limit transformers switches to FACTORISED buckets [2^(n//2), 2^(n - n//2)] (different rotation
        # shape, RNG consumption and ids): 128 buckets at chunk 64, 256 at chunk 128, i.e. T >= 8192 / 32768.  Not built.
        limit = 2 * max(int((self.max_position_embeddings // self.bucket_size) ** 0.5), self.bucket_size)
        if nb > limit:
            raise NotImplementedError(f"HFLSHSelfAttention: {nb} buckets exceed transformers' limit of {limit} for chunk length "
                                      f"{self.bucket_size}; the factorised bucket hashing it switches to (hf:788-793) is not built")
        rot = torch.randn((self.heads, d // self.heads, self.n_hashes, nb // 2), dtype=torch.float32, device=x.device)  # hf:717-719
        if self.rot_override is not None:
            rot = self.rot_override.to(device=x.device, dtype=torch.float32)
        # hf:740-747 sends padded tokens to an extra bucket if the mask masks anything (a host sync in the reference).  Here the
        # extra bucket is used whenever a mask is supplied: empty, it changes nothing but the round stride of the ids (same sort
        # order, same attention), so there is no host decision to read - nothing a CUDA-graph capture could freeze wrongly.
        pad_bucket = attention_mask is not None
        cfg = dict(heads=self.heads, n_hashes=self.n_hashes, bucket_size=self.bucket_size, n_buckets=nb, pad_bucket=pad_bucket,
                   spec=LSHSpec.huggingface(d // self.heads, self.causal), eps=1e-5)
        return self._run(x, norm, self.query_key.weight, self.value.weight, None, None, rot, attention_mask, cfg)
