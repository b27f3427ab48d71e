"""Operand-rounding mode of the LSH-attention oracle (TEST INFRASTRUCTURE ONLY).

``oracle.lsh_core`` restates the reference arithmetic in exact fp32 / fp64.  The CUDA path computes the same
quantities with bf16 OPERANDS and fp32 accumulation (BASELINE.json north_star), so some intermediates are rounded
to bf16 where they are stored as the operand of the next tensor-core contraction.  A comparison of the kernels with
the exact oracle therefore measures storage rounding (2^-9/sqrt(3) = 1.1e-3 relative per bf16 hop), not kernel
correctness.  This file restates the SAME algorithm (rp R4-R11 / hf:563-645,801-906 and its analytic gradient) with
the roundings applied at exactly the boundaries the kernels have, so that the 1e-3 criterion of north_star can be
asserted as a fact ("same operands pre-rounded to bf16, oracle in fp32").

With ``round_operands=False`` nothing is rounded and

* ``forward``  must equal ``lsh_core.lsh_attention`` (checked in tests/test_oracle.py), and
* ``backward`` must equal autograd through ``lsh_core.lsh_attention`` (checked there too, fp64, 1e-9):
  the analytic gradient is an independent restatement, not a transcription of the kernels.

Rounding boundaries mirrored (reformer_tts_b200/csrc):
  forward   P = exp2(score - stabiliser) -> bf16 (A operand of O = P V, lsh_attn_fwd.cu), stabiliser = the
            Cauchy-Schwarz bound |q| * scale * log2(e) * 1.001 (rows whose bound reaches 60: true row maximum);
            row sum as the kernel forms it (``KERNEL_SUM_ROUNDED``);
            per-round o -> bf16 (o_rounds); merged out -> bf16.
  backward  Pt = exp2(score - L) -> bf16 (A operand of dV), dS' = Pt (dP - delta) scale / |k| -> bf16 (operand of
            dQ and G), the per-round partial sums (dv; G-Jacobian + main-chunk dq; look-back dq) -> bf16, the sum
            over rounds in fp32 -> bf16 dqk | dv (lsh_attn_bwd.cu, lsh_grad_reduce_kernel).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

from .lsh_core import KEYNORM_L2, MASK_KEY_ONLY, LSHSpec

LOG2E = 1.4426950408889634
LN2 = 0.6931471805599453
EXACT_BOUND = 60.0        # lsh_attn_fwd.cu kExactBound
# True: the forward kernel's softmax normaliser is the sum of the bf16-ROUNDED P (it comes out of the same tensor-core
# contraction as O = P V: a ones block appended to V); False: the fp32 sum of the unrounded terms.
KERNEL_SUM_ROUNDED = True


def bf16r(x: torch.Tensor) -> torch.Tensor:
    """Round to bf16 (nearest even) and return in the input dtype."""
    return x.to(torch.bfloat16).to(x.dtype)


def _maybe(x: torch.Tensor, on: bool) -> torch.Tensor:
    return bf16r(x) if on else x


def _with_prev(x: torch.Tensor) -> torch.Tensor:
    """[N, C, bs, ...] -> [N, C, 2*bs, ...]: chunk c followed by chunk c-1 (rp R6; wraps)."""
    return torch.cat([x, torch.roll(x, shifts=1, dims=1)], dim=2)


def _key_inv_norm(x: torch.Tensor, spec: LSHSpec) -> torch.Tensor:
    """1 / |k| of the key normalisation (rp R5 / hf:1042-1056), per row."""
    if spec.key_norm == KEYNORM_L2:
        return 1.0 / x.norm(dim=-1).clamp_min(1e-12)
    return torch.rsqrt((x * x).mean(dim=-1) + 1e-6) / math.sqrt(x.shape[-1])


def _geometry(qk, v, sticker, bucket, n_rounds, spec: LSHSpec, mask):
    """Gathered chunks and the three masks every stage needs."""
    n, t, dh = qk.shape
    st = sticker % t
    chunks = n_rounds * t // bucket
    idx = st.unsqueeze(-1).expand(-1, -1, dh)
    bq = qk.gather(1, idx).reshape(n, chunks, bucket, dh)            # rp R4
    bv = v.gather(1, idx).reshape(n, chunks, bucket, dh)
    q_pos = st.reshape(n, chunks, bucket)
    k_pos = _with_prev(q_pos)
    masked = torch.zeros(n, chunks, bucket, 2 * bucket, dtype=torch.bool)
    if mask is not None:
        mq = mask.gather(1, st).reshape(n, chunks, bucket)
        mk = _with_prev(mq)
        keep = mk[:, :, None, :] if spec.mask_mode == MASK_KEY_ONLY else mq[:, :, :, None] & mk[:, :, None, :]
        masked = ~keep
    if spec.causal:
        masked = masked | (q_pos[:, :, :, None] < k_pos[:, :, None, :])
    is_self = q_pos[:, :, :, None] == k_pos[:, :, None, :]
    return bq, bv, q_pos, masked, is_self, chunks


def forward(qk: torch.Tensor, v: torch.Tensor, sticker: torch.Tensor, undo: torch.Tensor, bucket: int, n_rounds: int,
            spec: LSHSpec, mask: Optional[torch.Tensor] = None, round_operands: bool = True,
            sum_rounded: Optional[bool] = None) -> Dict[str, torch.Tensor]:
    """qk, v [N,T,dh] (bf16-representable values when rounding); returns o_rounds [N,R,T,dh], lse_rounds [N,R,T],
    out [N,T,dh], lse [N,T] with the kernels' bf16 storage roundings applied to o_rounds and out."""
    n, t, dh = qk.shape
    bq, bv, q_pos, masked, is_self, chunks = _geometry(qk, v, sticker, bucket, n_rounds, spec, mask)
    inv = _key_inv_norm(bq, spec)                                            # [N,C,bs]
    ks = inv * (spec.score_scale * LOG2E)                                    # score -> log2 units, per key
    dots = torch.einsum("ncid,ncjd->ncij", bq, _with_prev(bq)) * _with_prev(ks)[:, :, None, :]
    mv = max(spec.mask_value * LOG2E, -3.0e38)
    sv = spec.self_value * LOG2E
    ref_scores = torch.where(masked, torch.full_like(dots, mv), dots)
    ref_scores = torch.where(is_self, torch.full_like(dots, sv), ref_scores)   # rp R8: self overrides the other fills
    row_max = ref_scores.max(dim=-1).values
    if round_operands:
        # single-pass stabiliser of the kernel; rows whose bound could underflow a visible key use the true maximum
        bound = (spec.score_scale * LOG2E) ** 2 / ks * 1.001
        m = torch.where(bound >= EXACT_BOUND, row_max, bound)
    else:
        m = row_max
    p = torch.exp2(ref_scores - m[..., None])
    pr = _maybe(p, round_operands)
    row_sum = (pr if (KERNEL_SUM_ROUNDED if sum_rounded is None else sum_rounded) else p).sum(dim=-1)
    lonely = row_sum <= 0          # every visible term underflowed: the row sees only itself (kernel: analytic result)
    if bool(lonely.any()):
        n_self = is_self.sum(dim=-1).to(dots.dtype)
        uniform = is_self.to(dots.dtype) / n_self.clamp_min(1)[..., None]
        pr = torch.where(lonely[..., None], uniform, pr)
        row_sum = torch.where(lonely, torch.ones_like(row_sum), row_sum)
        m = torch.where(lonely, sv + torch.log2(n_self.clamp_min(1)), m)
    so = torch.einsum("ncij,ncjd->ncid", pr, _with_prev(bv)) / row_sum[..., None]
    slse = (m + torch.log2(row_sum)) * LN2
    so, slse = so.reshape(n, -1, dh), slse.reshape(n, -1)
    o = _maybe(so.gather(1, undo.unsqueeze(-1).expand(-1, -1, dh)).reshape(n, n_rounds, t, dh), round_operands)   # rp R10
    lse_r = slse.gather(1, undo).reshape(n, n_rounds, t)
    lse = torch.logsumexp(lse_r, dim=1)
    w = torch.exp(lse_r - lse[:, None, :])                                                                       # rp R11
    out = _maybe((o * w.unsqueeze(-1)).sum(dim=1), round_operands)
    return {"o_rounds": o, "lse_rounds": lse_r, "out": out, "lse": lse}


def backward(qk: torch.Tensor, v: torch.Tensor, sticker: torch.Tensor, undo: torch.Tensor, bucket: int, n_rounds: int,
             spec: LSHSpec, mask: Optional[torch.Tensor], dout: torch.Tensor, out: torch.Tensor, lse: torch.Tensor,
             round_operands: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """Analytic gradient of (hash-rounds merged) LSH attention w.r.t. qk and v, [N,T,dh] each.

    Across rounds the layer is ONE softmax with normaliser L = logsumexp_r lse_r, so with Pt = exp(s - L),
    delta_i = <dout_i, out_i>:  dV_j += Pt_ij dout_i,  dS_ij = Pt_ij (<dout_i, v_j> - delta_i) (zero where a constant
    was written: masked or self), dQ_i += dS_ij scale khat_j, dKhat_j += dS_ij scale q_i, and the key gradient passes
    the normalisation Jacobian dx = (dKhat - khat <khat, dKhat>) / |x|."""
    n, t, dh = qk.shape
    bq, bv, q_pos, masked, is_self, chunks = _geometry(qk, v, sticker, bucket, n_rounds, spec, mask)
    st = sticker % t
    gather_q = lambda a: a.gather(1, st).reshape(n, chunks, bucket)      # per-token [N,T] -> per sorted slot
    l2 = gather_q(lse) * LOG2E
    delta = gather_q((dout * out).sum(dim=-1))
    bdo = dout.gather(1, st.unsqueeze(-1).expand(-1, -1, dh)).reshape(n, chunks, bucket, dh)
    inv = _key_inv_norm(bq, spec)
    cs = _with_prev(inv * (spec.score_scale * LOG2E))[:, :, None, :]
    gs = _with_prev(inv * spec.score_scale)[:, :, None, :]
    xk, vk = _with_prev(bq), _with_prev(bv)
    s = torch.einsum("ncid,ncjd->ncij", bq, xk)
    pt = torch.exp2(s * cs - l2[..., None])
    pt = torch.where(masked, torch.zeros_like(pt), pt)
    sv = spec.self_value * LOG2E
    pt = torch.where(is_self, torch.exp2(sv - l2)[..., None].expand_as(pt), pt)
    dp = torch.einsum("ncid,ncjd->ncij", bdo, vk)
    ds = pt * ((dp - delta[..., None]) * gs)
    ds = torch.where(is_self | masked, torch.zeros_like(ds), ds)
    pt, ds = _maybe(pt, round_operands), _maybe(ds, round_operands)
    # key-side sums: a key of chunk c is seen by the queries of chunk c (first half of the window) and of chunk c+1 (second half)
    def key_side(w, rows):      # w [N,C,bs_q,2bs_k], rows [N,C,bs_q,dh] -> [N,C,bs_k,dh]
        both = torch.einsum("ncij,ncid->ncjd", w, rows)
        return both[:, :, :bucket] + torch.roll(both[:, :, bucket:], shifts=-1, dims=1)
    dv_s = key_side(pt, bdo)
    g = key_side(ds, bq)                                                # = dKhat * scale / |x|
    dx = g - bq * (inv * inv * (bq * g).sum(dim=-1))[..., None]         # normalisation Jacobian
    dq_main = torch.einsum("ncij,ncjd->ncid", ds[..., :bucket], bq)     # keys of the query's own chunk
    dq_lb = torch.einsum("ncij,ncjd->ncid", ds[..., bucket:], torch.roll(bq, shifts=1, dims=1))
    if round_operands:
        # per-round partial tensors of the backward kernel (bf16): a tile is 128 sorted slots = its key chunks; a query chunk that
        # is the FIRST chunk of a tile gets its look-back part from the previous tile (dq_b), any other chunk is complete in one tile
        per_tile = 128 // bucket
        first_in_tile = (torch.arange(chunks) % per_tile == 0).view(1, chunks, 1, 1)
        dqk_s = torch.where(first_in_tile, bf16r(dq_main + dx) + bf16r(dq_lb), bf16r(dq_main + dq_lb + dx))
        dv_s = bf16r(dv_s)
    else:
        dqk_s = dq_main + dq_lb + dx
    unsort = lambda a: a.reshape(n, -1, dh).gather(1, undo.unsqueeze(-1).expand(-1, -1, dh)).reshape(n, n_rounds, t, dh).sum(dim=1)
    return _maybe(unsort(dqk_s), round_operands), _maybe(unsort(dv_s), round_operands)
