"""Shared helpers of the parity tests."""
import torch

# Tolerances (BASELINE.json north_star: bucket ids / permutations bit-exact; activations and gradients within 1e-3
# relative with fp32 accumulation and bf16 operands).
TOL_FP32 = 1e-3     # relative L2 error of anything accumulated and STORED in fp32 (lse, GEMM fp32 outputs, LayerNorm grads, weight grads)
# A tensor that is itself STORED as a bf16 operand of the next kernel (qk|v, P, o_rounds, merged out, dqk/dv, hidden h) carries
# the storage rounding of bf16 (8 significant bits): relative L2 error 2^-9/sqrt(3) = 1.1e-3 per rounding.  Such tensors are
# compared with the oracle at 3e-3 (two to three roundings in the chain: P, the stored result, and the bf16 gradient operand).
TOL_BF16_STORED = 3e-3
# Gradients of the attention core go through four bf16 hops (P~, dS, the per-round partial sums that are written to HBM in bf16 and
# summed over rounds in fp32, and the final bf16 operand of the projection-gradient GEMM): 4e-3 (measured 2.6e-3 .. 3.1e-3 on N(0,1) data).
TOL_BF16_GRAD = 4e-3


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def to_bh(x: torch.Tensor, heads: int) -> torch.Tensor:
    """token-major [B,T,H*dh] -> [B*H, T, dh] fp32 on CPU (the oracle's layout)."""
    b, t, c = x.shape
    return x.detach().float().view(b, t, heads, c // heads).transpose(1, 2).reshape(b * heads, t, c // heads).cpu()
