"""Torch-tensor front end of the C ABI: checks, output allocation, stream plumbing.  No arithmetic here.

Every function launches on ``torch.cuda.current_stream()`` and raises ``RuntimeError`` on a non-zero
return from the library.  PyTorch is only the allocator / stream provider.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib

KEYNORM_L2, KEYNORM_RMS = 0, 1
MASK_QUERY_AND_KEY, MASK_KEY_ONLY = 0, 1
from .residual import GradAccumRequest, ResidualRequest, grad_sink  # noqa: F401  (re-exported: layers call ops.ResidualRequest.take)

EPI_BIAS, EPI_RELU, EPI_GATE, EPI_OUT_BF16, EPI_ATOMIC, EPI_COLSUM, EPI_RESID_ADD, EPI_RESID_SUB = 1, 2, 4, 8, 16, 32, 64, 128


@dataclass(frozen=True)
class LSHSpec:
    """Constants separating the two implementations ref:reformer_tts/model/reformer.py:198-213 can select."""
    score_scale: float
    key_norm: int
    mask_value: float
    self_value: float
    mask_mode: int
    causal: bool

    @staticmethod
    def reformer_pytorch(dh: int, causal: bool) -> "LSHSpec":      # rp R5, R7, R8
        return LSHSpec(dh ** -0.5, KEYNORM_L2, -torch.finfo(torch.float32).max, -5e4, MASK_QUERY_AND_KEY, causal)

    @staticmethod
    def huggingface(dh: int, causal: bool) -> "LSHSpec":           # hf:430-431,914-922,1042-1056
        return LSHSpec(1.0, KEYNORM_RMS, -1e9, -1e5, MASK_KEY_ONLY, causal)

    def struct(self) -> _lib.LSHSpecStruct:
        return _lib.LSHSpecStruct(self.score_scale, self.key_norm, self.mask_value, self.self_value, self.mask_mode,
                                  int(self.causal))


# ---------------------------------------------------------------------------------------------------------------------
# Launch accounting and optional per-kernel CUDA-event timing (bench.py: gpu_launches, roofline.achieved)
_LAUNCHES = 0
_TIMER = None


class KernelTimer:
    """Records a CUDA-event pair on the launching stream around selected kernels; ``summary()`` after a synchronize."""

    def __init__(self, only=None):
        self.only = None if only is None else tuple(only)
        self.records = []      # (name, start, end)

    def wants(self, name: str) -> bool:
        return self.only is None or name.split("[")[0] in self.only

    def summary(self):
        out = {}
        for name, a, b in self.records:
            e = out.setdefault(name, {"count": 0, "total_ms": 0.0})
            e["count"] += 1
            e["total_ms"] += a.elapsed_time(b)
        for e in out.values():
            e["avg_ms"] = e["total_ms"] / e["count"]
        return out


def set_kernel_timer(timer: Optional["KernelTimer"]):
    global _TIMER
    _TIMER = timer


def launch_count() -> int:
    """Number of kernels of libreformer_b200.so launched so far by this process."""
    return _LAUNCHES


def _launch(tag: str, name: str, *args):
    global _LAUNCHES
    _LAUNCHES += 1
    timer = _TIMER
    if timer is not None and timer.wants(tag):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # In an eager pass the host is the slow side: the GPU would reach event `a` idle and then wait for the launch to be enqueued,
        # and that host latency (5-10 us) would be counted as kernel time.  A short device-side spin keeps the stream busy until the
        # event, the kernel and the closing event are all queued.
        torch.cuda._sleep(40000)
        a.record()
        _lib.call(name, *args)
        b.record()
        timer.records.append((tag, a, b))
    else:
        _lib.call(name, *args)


def _tag(name: str, scope: dict) -> str:
    """Kernel name plus the shape key the benchmark groups by (sequence length for the LSH kernels, MxNxK for GEMMs)."""
    if name.startswith("lsh_") and "t" in scope and isinstance(scope["t"], int):
        return f"{name}[T={scope['t']}]"
    if name == "gemm_bf16":
        return f"gemm_bf16[{scope['m']}x{scope['n']}x{scope['k']}]"
    return name


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check(t: torch.Tensor, dtype, name: str, inner_contig: bool = True):
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected {dtype}, got {t.dtype}")
    if inner_contig and t.stride(-1) != 1:
        raise RuntimeError(f"{name}: innermost dimension must be contiguous")


def _token_major(t: torch.Tensor, name: str) -> int:
    """[B, T, C] bf16 view with contiguous channels and B-stride == T * ld; returns ld."""
    _check(t, torch.bfloat16, name)
    b, tt, _ = t.shape
    ld = t.stride(1)
    if b > 1 and t.stride(0) != tt * ld:
        raise RuntimeError(f"{name}: batch stride must equal T * token stride")
    return ld


def lsh_hash(qk: torch.Tensor, rot: torch.Tensor, n_heads: int, n_rounds: int, n_buckets: int,
             pad_mask: Optional[torch.Tensor] = None, use_pad_bucket: bool = False, return_sumsq: bool = False):
    """qk bf16 [B,T,H*64]; rot fp32 [1|H, 64, R, nb/2]  ->  buckets int32 [B,H,R*T] (and |qk row|^2 fp32 [B,H,T])."""
    ld = _token_major(qk, "qk")
    b, t, c = qk.shape
    dh = c // n_heads
    _check(rot, torch.float32, "rot")
    rot = rot.contiguous()
    assert rot.shape[1:] == (dh, n_rounds, n_buckets // 2), f"rot shape {tuple(rot.shape)}"
    if pad_mask is not None:
        _check(pad_mask, torch.uint8, "pad_mask")
        pad_mask = pad_mask.contiguous()
    out = torch.empty((b, n_heads, n_rounds * t), dtype=torch.int32, device=qk.device)
    sumsq = torch.empty((b, n_heads, t), dtype=torch.float32, device=qk.device) if return_sumsq else None
    if _lib.load().rtts_lsh_hash_tc_supported(t, dh, n_rounds, n_buckets):
        # tensor-pipe path: exact products of the bf16 rows with the three-way bf16 split of the fp32 rotations
        ws = torch.empty(_lib.load().rtts_lsh_hash_tc_workspace_bytes(rot.shape[0], n_rounds, n_buckets) // 2, dtype=torch.bfloat16, device=qk.device)
        _launch(_tag("lsh_hash", locals()), "rtts_lsh_hash_tc", _ptr(qk), ld, _ptr(rot), rot.shape[0], _ptr(pad_mask), int(use_pad_bucket),
                _ptr(out), _ptr(sumsq), _ptr(ws), b, t, n_heads, dh, n_rounds, n_buckets, _stream())
    else:
        _launch(_tag("lsh_hash", locals()), "rtts_lsh_hash", _ptr(qk), ld, _ptr(rot), rot.shape[0], _ptr(pad_mask), int(use_pad_bucket), _ptr(out),
                _ptr(sumsq), b, t, n_heads, dh, n_rounds, n_buckets, _stream())
    return (out, sumsq) if return_sumsq else out


def lsh_sumsq(qk: torch.Tensor, n_heads: int) -> torch.Tensor:
    """|qk[b,t,h,:]|^2 -> fp32 [B,H,T]."""
    ld = _token_major(qk, "qk")
    b, t, c = qk.shape
    sumsq = torch.empty((b, n_heads, t), dtype=torch.float32, device=qk.device)
    _launch(_tag("lsh_sumsq", locals()), "rtts_lsh_sumsq", _ptr(qk), ld, _ptr(sumsq), b, t, n_heads, c // n_heads, _stream())
    return sumsq


def lsh_sort(buckets: torch.Tensor, seq_len: int, n_rounds: int, ids_per_round: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """buckets int32 [..., R*T] -> (sticker, undo) int32, same shape."""
    _check(buckets, torch.int32, "buckets")
    buckets = buckets.contiguous()
    rows = buckets.numel() // (n_rounds * seq_len)
    sticker, undo = torch.empty_like(buckets), torch.empty_like(buckets)
    _launch(_tag("lsh_sort", locals()), "rtts_lsh_sort", _ptr(buckets), _ptr(sticker), _ptr(undo), rows, seq_len, n_rounds, ids_per_round, _stream())
    return sticker, undo


def lsh_attn_fwd(qk: torch.Tensor, v: torch.Tensor, sticker: torch.Tensor, mask: Optional[torch.Tensor], spec: LSHSpec,
                 n_heads: int, n_rounds: int, bucket: int, sumsq: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> o_rounds bf16 [B,H,R,T,64], lse_rounds fp32 [B,H,R,T] (both already unsorted)."""
    ld = _token_major(qk, "qk")
    if _token_major(v, "v") != ld:
        raise RuntimeError("qk and v must share the token stride")
    b, t, c = qk.shape
    dh = c // n_heads
    _check(sticker, torch.int32, "sticker")
    assert sticker.is_contiguous() and sticker.numel() == b * n_heads * n_rounds * t
    if mask is not None:
        _check(mask, torch.uint8, "mask")
        mask = mask.contiguous()
    if sumsq is None:
        sumsq = lsh_sumsq(qk, n_heads)
    _check(sumsq, torch.float32, "sumsq")
    assert sumsq.is_contiguous() and sumsq.numel() == b * n_heads * t
    o = torch.empty((b, n_heads, n_rounds, t, dh), dtype=torch.bfloat16, device=qk.device)
    lse = torch.empty((b, n_heads, n_rounds, t), dtype=torch.float32, device=qk.device)
    st = spec.struct()
    _launch(_tag("lsh_attn_fwd", locals()), "rtts_lsh_attn_fwd", _ptr(qk), _ptr(v), ld, _ptr(sticker), _ptr(sumsq), _ptr(mask), ctypes.byref(st), _ptr(o), _ptr(lse),
              b, t, n_heads, dh, n_rounds, bucket, _stream())
    return o, lse


def lsh_merge_fwd(o_rounds: torch.Tensor, lse_rounds: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> out bf16 [B,T,H*64], lse fp32 [B,H,T]."""
    _check(o_rounds, torch.bfloat16, "o_rounds")
    _check(lse_rounds, torch.float32, "lse_rounds")
    b, h, r, t, dh = o_rounds.shape
    assert o_rounds.is_contiguous() and lse_rounds.is_contiguous()
    out = torch.empty((b, t, h * dh), dtype=torch.bfloat16, device=o_rounds.device)
    lse = torch.empty((b, h, t), dtype=torch.float32, device=o_rounds.device)
    _launch(_tag("lsh_merge_fwd", locals()), "rtts_lsh_merge_fwd", _ptr(o_rounds), _ptr(lse_rounds), _ptr(out), h * dh, _ptr(lse), b, t, h, dh, r, _stream())
    return out, lse


def lsh_delta(dout: torch.Tensor, out: torch.Tensor, n_heads: int) -> torch.Tensor:
    ld = _token_major(dout, "dout")
    if _token_major(out, "out") != ld:
        raise RuntimeError("dout and out must share the token stride")
    b, t, c = dout.shape
    delta = torch.empty((b, n_heads, t), dtype=torch.float32, device=dout.device)
    _launch(_tag("lsh_delta", locals()), "rtts_lsh_delta", _ptr(dout), _ptr(out), ld, _ptr(delta), b, t, n_heads, c // n_heads, _stream())
    return delta


def lsh_attn_bwd(qk, v, sticker, undo, mask, spec: LSHSpec, dout, lse, delta, n_heads: int, n_rounds: int, bucket: int,
                 out_dqk: Optional[torch.Tensor] = None, out_dv: Optional[torch.Tensor] = None, sumsq: Optional[torch.Tensor] = None):
    """Backward of lsh_attn_fwd + lsh_merge_fwd (scores recomputed in-kernel) -> dqk, dv bf16 [B,T,H*64]."""
    ld = _token_major(qk, "qk")
    if _token_major(v, "v") != ld:
        raise RuntimeError("qk and v must share the token stride")
    ld_do = _token_major(dout, "dout")
    b, t, c = qk.shape
    dh = c // n_heads
    _check(lse, torch.float32, "lse")
    _check(delta, torch.float32, "delta")
    _check(undo, torch.int32, "undo")
    assert lse.is_contiguous() and delta.is_contiguous() and sticker.is_contiguous() and undo.is_contiguous()
    if sumsq is None:
        sumsq = lsh_sumsq(qk, n_heads)
    _check(sumsq, torch.float32, "sumsq")
    # bf16 per-round partials [3, B,H,R,T,dh]: dqk_main, dq_b, dv (summed over rounds in fp32 by the reduce kernel)
    part = torch.empty((3, b, n_heads, n_rounds, t, dh), dtype=torch.bfloat16, device=qk.device)
    st = spec.struct()
    _launch(_tag("lsh_attn_bwd", locals()), "rtts_lsh_attn_bwd", _ptr(qk), _ptr(v), ld, _ptr(sticker), _ptr(sumsq), _ptr(mask), ctypes.byref(st), _ptr(dout), ld_do,
              _ptr(lse), _ptr(delta), _ptr(part[0]), _ptr(part[1]), _ptr(part[2]), b, t, n_heads, dh, n_rounds,
              bucket, _stream())
    dqk = torch.empty((b, t, c), dtype=torch.bfloat16, device=qk.device) if out_dqk is None else out_dqk
    dv = torch.empty((b, t, c), dtype=torch.bfloat16, device=qk.device) if out_dv is None else out_dv
    ld_out = _token_major(dqk, "out_dqk")
    if _token_major(dv, "out_dv") != ld_out:
        raise RuntimeError("out_dqk and out_dv must share the token stride")
    _launch(_tag("lsh_grad_reduce", locals()), "rtts_lsh_grad_reduce", _ptr(part[0]), _ptr(part[1]), _ptr(part[2]), _ptr(undo), _ptr(dqk),
              _ptr(dv), ld_out, b, t, n_heads, dh, n_rounds, bucket, _stream())
    return dqk, dv


def layernorm_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5, save_stats: bool = True):
    """x fp32 [..., dim] -> (y bf16, mean, rstd)."""
    _check(x, torch.float32, "x")
    x = x.contiguous()
    dim = x.shape[-1]
    rows = x.numel() // dim
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    mean = torch.empty(rows, dtype=torch.float32, device=x.device) if save_stats else None
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device) if save_stats else None
    _launch(_tag("layernorm_fwd", locals()), "rtts_layernorm_fwd", _ptr(x), _ptr(gamma.contiguous()), _ptr(beta.contiguous()), _ptr(y), _ptr(mean), _ptr(rstd),
              rows, dim, eps, _stream())
    return y, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta, accumulate_request: bool = False):
    """dy, x fp32 [..., dim] -> dx fp32; dgamma / dbeta (fp32 [dim]) are accumulated in place.  With ``accumulate_request`` a pending
    ``GradAccumRequest`` of the same shape is taken: dx = request.base + (LayerNorm backward), one pass instead of two."""
    _check(dy, torch.float32, "dy")
    dy, x = dy.contiguous(), x.contiguous()
    dim = x.shape[-1]
    dx = torch.empty_like(x)
    add = GradAccumRequest.take(x.numel(), x.device) if accumulate_request else None
    _launch(_tag("layernorm_bwd", locals()), "rtts_layernorm_bwd_acc", _ptr(dy), _ptr(x), _ptr(gamma.contiguous()), _ptr(mean), _ptr(rstd), _ptr(add),
              _ptr(dx), _ptr(dgamma), _ptr(dbeta), x.numel() // dim, dim, _stream())
    return dx


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_mn_major: bool = False, b_mn_major: bool = False, bias=None, relu=False,
         gate=None, out_dtype=torch.float32, out: Optional[torch.Tensor] = None, accumulate: bool = False,
         colsum: Optional[torch.Tensor] = None, split_k: int = 1, resid: Optional[torch.Tensor] = None, resid_sub: bool = False,
         keep_mask: Optional[torch.Tensor] = None, keep_scale: float = 1.0) -> torch.Tensor:
    """C = epilogue(A . B^T) with bf16 operands.  A: [M,K] (or stored [K,M] when a_mn_major); B: [N,K] (or [K,N]).
    ``resid`` (fp32 [M,N]): C = resid + result, or resid - result with ``resid_sub`` (fp32 output only).
    ``keep_mask`` (uint8 [M,N], 1 = keep): inverted dropout on the result, ``keep * keep_scale * result`` (fp32 output only)."""
    _check(a, torch.bfloat16, "a")
    _check(b, torch.bfloat16, "b")
    assert a.dim() == 2 and b.dim() == 2
    m, k = (a.shape[1], a.shape[0]) if a_mn_major else a.shape
    n, kb = (b.shape[1], b.shape[0]) if b_mn_major else b.shape
    assert k == kb, f"K mismatch {k} vs {kb}"
    flags = 0
    if bias is not None:
        _check(bias, torch.float32, "bias")
        flags |= EPI_BIAS
    if relu:
        flags |= EPI_RELU
    if gate is not None:
        _check(gate, torch.bfloat16, "gate")
        flags |= EPI_GATE
    if colsum is not None:
        _check(colsum, torch.float32, "colsum")
        flags |= EPI_COLSUM
    if resid is not None:
        _check(resid, torch.float32, "resid")
        assert gate is None and not accumulate and out_dtype == torch.float32 and resid.dim() == 2 and resid.stride(1) == 1
        flags |= EPI_RESID_SUB if resid_sub else EPI_RESID_ADD
        gate = resid
    if accumulate:
        assert out is not None and out.dtype == torch.float32
        flags |= EPI_ATOMIC
    elif out_dtype == torch.bfloat16:
        flags |= EPI_OUT_BF16
    if out is None:
        out = torch.empty((m, n), dtype=out_dtype, device=a.device)
    assert out.shape == (m, n) and out.stride(1) == 1
    if keep_mask is not None:
        _check(keep_mask, torch.uint8, "keep_mask")
        assert keep_mask.shape == (m, n) and keep_mask.stride(1) == 1 and out.dtype == torch.float32 and not accumulate and colsum is None
        _launch(_tag("gemm_bf16", locals()), "rtts_gemm_bf16_dropout", _ptr(a), a.stride(0), int(a_mn_major), _ptr(b), b.stride(0), int(b_mn_major),
                _ptr(out), out.stride(0), _ptr(bias), _ptr(gate), 0 if gate is None else gate.stride(0), _ptr(colsum), _ptr(keep_mask),
                keep_mask.stride(0), float(keep_scale), m, n, k, flags, split_k, _stream())
        return out
    _launch(_tag("gemm_bf16", locals()), "rtts_gemm_bf16", _ptr(a), a.stride(0), int(a_mn_major), _ptr(b), b.stride(0), int(b_mn_major), _ptr(out),
              out.stride(0), _ptr(bias), _ptr(gate), 0 if gate is None else gate.stride(0), _ptr(colsum), m, n, k, flags, split_k,
              _stream())
    return out


def xattn_supported(t: int, s: int, dh: int) -> bool:
    """Shapes the dense cross-attention kernels take (rtts_xattn_fwd / _bwd)."""
    return dh == 64 and t % 128 == 0 and s % 32 == 0 and 32 <= s <= 256


def _xattn_args(q, k, v, keep, n_heads):
    _check(q, torch.bfloat16, "q"); _check(k, torch.bfloat16, "k"); _check(v, torch.bfloat16, "v")
    b, t, d = q.shape
    s = k.shape[1]
    assert q.stride(2) == 1 and q.stride(0) == t * q.stride(1) and k.stride(2) == 1 and v.stride() == k.stride() and k.stride(0) == s * k.stride(1)
    if keep is not None:
        _check(keep, torch.uint8, "keep")
        assert keep.is_contiguous() and keep.shape == (b, s)
    return b, t, s, d // n_heads


def xattn_fwd(q, k, v, keep, n_heads: int, scale: float, p_drop: float = 0.0, seed: Optional[torch.Tensor] = None):
    """softmax(q k^T scale) v per head with key padding (keep uint8 [B,S], 1 = valid) and probability dropout keyed by the device
    word ``seed`` (int64 [1]).  q [B,T,D], k / v [B,S,D] views (bf16, unit inner stride) -> out bf16 [B,T,D], lse fp32 [B,H,T]."""
    b, t, s, dh = _xattn_args(q, k, v, keep, n_heads)
    out = torch.empty((b, t, q.shape[2]), dtype=torch.bfloat16, device=q.device)
    lse = torch.empty((b, n_heads, t), dtype=torch.float32, device=q.device)
    _launch("xattn_fwd", "rtts_xattn_fwd", _ptr(q), q.stride(1), _ptr(k), _ptr(v), k.stride(1), _ptr(keep), float(scale), float(p_drop), _ptr(seed),
            _ptr(out), out.stride(1), _ptr(lse), b, t, s, n_heads, dh, _stream())
    return out, lse


def xattn_bwd(q, k, v, keep, n_heads: int, scale: float, p_drop: float, seed, dout, lse, delta):
    """Backward of xattn_fwd -> dq bf16 [B,T,D], dkv fp32 [B,S,2D] (dk | dv)."""
    b, t, s, dh = _xattn_args(q, k, v, keep, n_heads)
    _check(dout, torch.bfloat16, "dout")
    d = q.shape[2]
    assert dout.shape == q.shape and dout.stride(2) == 1 and dout.stride(0) == t * dout.stride(1)
    dq = torch.empty((b, t, d), dtype=torch.bfloat16, device=q.device)
    dkv = torch.zeros((b, s, 2 * d), dtype=torch.float32, device=q.device)
    _launch("xattn_bwd", "rtts_xattn_bwd", _ptr(q), q.stride(1), _ptr(k), _ptr(v), k.stride(1), _ptr(keep), float(scale), float(p_drop), _ptr(seed),
            _ptr(dout), dout.stride(1), _ptr(lse), _ptr(delta), _ptr(dq), dq.stride(1), _ptr(dkv), _ptr(dkv[..., d:]), dkv.stride(1), b, t, s, n_heads, dh,
            _stream())
    return dq, dkv


def colsum_bf16(x: torch.Tensor, colsum: torch.Tensor) -> torch.Tensor:
    """colsum (fp32 [cols]) += column sums of the bf16 matrix x [rows, cols] (rows may be strided): bias gradients."""
    _check(x, torch.bfloat16, "x")
    _check(colsum, torch.float32, "colsum")
    assert x.dim() == 2 and x.stride(1) == 1 and colsum.is_contiguous() and colsum.numel() == x.shape[1]
    _launch("colsum_bf16", "rtts_colsum_bf16", _ptr(x), x.stride(0), _ptr(colsum), x.shape[0], x.shape[1], _stream())
    return colsum


def cast_bf16_colsum(x: torch.Tensor, colsum: Optional[torch.Tensor] = None, keep_mask: Optional[torch.Tensor] = None,
                     keep_scale: float = 1.0) -> torch.Tensor:
    """fp32 [rows, cols] -> bf16 copy; colsum (fp32 [cols]) += column sums.  With ``keep_mask`` (uint8, x's shape) the inverted
    dropout ``keep * keep_scale * x`` is applied first (backward of a GEMM whose epilogue applied the same mask)."""
    _check(x, torch.float32, "x")
    x = x.contiguous()
    cols = x.shape[-1]
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    if keep_mask is not None:
        _check(keep_mask, torch.uint8, "keep_mask")
        assert keep_mask.is_contiguous() and keep_mask.numel() == x.numel()
        _launch(_tag("cast_bf16_colsum", locals()), "rtts_cast_bf16_colsum_dropout", _ptr(x), _ptr(keep_mask), float(keep_scale), _ptr(y), _ptr(colsum),
                x.numel() // cols, cols, _stream())
        return y
    _launch(_tag("cast_bf16_colsum", locals()), "rtts_cast_bf16_colsum", _ptr(x), _ptr(y), _ptr(colsum), x.numel() // cols, cols, _stream())
    return y
