"""Shared helpers of the parity tests."""
import torch

# Tolerances (BASELINE.json north_star: bucket ids / permutations bit-exact; activations and gradients within 1e-3 relative
# with fp32 accumulation and bf16 operands).  Every floating-point result of the CUDA path is compared with the OPERAND-ROUNDED
# oracle (oracle/lsh_rounded.py, oracle/rounded.py: the reference arithmetic in fp32 with a bf16 rounding exactly where the
# kernels store a bf16 operand), so the 1e-3 below is the kernels' own error - accumulation order, exp2 / rsqrt approximations,
# the rare element that rounds the other way - and not bf16 storage rounding (2^-9/sqrt(3) = 1.1e-3 per hop, which a comparison
# with the exact fp32 oracle would measure instead; that number is printed for information, never asserted).
TOL = 1e-3
TOL_FP32 = TOL
TOL_BF16_STORED = TOL
TOL_BF16_GRAD = TOL


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def report(name: str, got: torch.Tensor, rounded: torch.Tensor, exact: torch.Tensor = None) -> float:
    """rel-L2 against the rounded oracle (returned, asserted by the caller); prints it next to the distance to the exact oracle."""
    r = rel_l2(got, rounded)
    msg = f"[parity] {name}: vs rounded oracle {r:.2e}"
    if exact is not None:
        msg += f", vs exact fp32 oracle {rel_l2(got, exact):.2e}"
    print(msg)
    return r


def bf16r(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(x.dtype)


def to_bh(x: torch.Tensor, heads: int) -> torch.Tensor:
    """token-major [B,T,H*dh] -> [B*H, T, dh] fp32 on CPU (the oracle's layout)."""
    b, t, c = x.shape
    return x.detach().float().view(b, t, heads, c // heads).transpose(1, 2).reshape(b * heads, t, c // heads).cpu()


def hf_inputs(dim, T, B, pad, seed):
    """Inputs of a gradient case, regenerated from the seed by the tests (the fixture stores only what transformers produced and
    a checksum of these tensors): weights, x, upstream gradient, padding mask."""
    g = torch.Generator().manual_seed(seed)
    wqk = torch.randn(dim, dim, generator=g) * dim ** -0.5
    wv = torch.randn(dim, dim, generator=g) * dim ** -0.5
    x = torch.randn(B, T, dim, generator=g)
    dy = torch.randn(B, T, dim, generator=g)
    mask = None
    if pad:
        mask = torch.ones(B, T, dtype=torch.bool)
        mask[0, T - T // 5:] = False
    checksum = float(wqk.double().sum() + 2 * wv.double().sum() + 3 * x.double().sum() + 5 * dy.double().sum())
    return wqk, wv, x, dy, mask, checksum
