"""One launch (or `iters`) of every kernel of the library at a given scale: the driver for compute-sanitizer (tiny shapes)
and for per-kernel ncu captures of the HBM-bound kernels (default-workload shapes).
  python tools/run_all_kernels.py tiny        B=1 T=256 H=2 R=2 (both bucket sizes, padding, causal and not)
  python tools/run_all_kernels.py full [n]    B=20 T=1024 H=8 R=8, n iterations (default 3)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reformer_tts_b200 import ops  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "tiny"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else (1 if mode == "tiny" else 3)
dev = "cuda"
torch.manual_seed(0)
shapes = [(1, 256, 2, 2, 64, 1, 40), (1, 256, 2, 2, 128, 0, 0), (2, 384, 2, 3, 64, 0, 30)] if mode == "tiny" else [(20, 1024, 8, 8, 64, 1, 0)]
for (B, T, H, R, bucket, causal, pad) in shapes:
    D = H * 64
    qkv = torch.randn(B, T, 2 * D, device=dev).bfloat16()
    qk, v = qkv[..., :D], qkv[..., D:]
    dout = torch.randn(B, T, D, device=dev).bfloat16()
    nb = T // bucket
    rot = torch.randn(1, 64, R, max(nb // 2, 1), device=dev)
    spec = ops.LSHSpec.reformer_pytorch(64, bool(causal))
    mask = None
    if pad:
        mask = torch.ones(B, T, dtype=torch.uint8, device=dev)
        mask[:, T - pad:] = 0
    for _ in range(iters):
        buckets, sumsq = ops.lsh_hash(qk, rot, H, R, nb, return_sumsq=True)
        sumsq2 = ops.lsh_sumsq(qk, H)
        sticker, undo = ops.lsh_sort(buckets, T, R, nb)
        o, lse_r = ops.lsh_attn_fwd(qk, v, sticker, mask, spec, H, R, bucket, sumsq=sumsq)
        out, lse = ops.lsh_merge_fwd(o, lse_r)
        delta = ops.lsh_delta(dout, out, H)
        dqk, dv = ops.lsh_attn_bwd(qk, v, sticker, undo, mask, spec, dout, lse, delta, H, R, bucket)
        # row-wise kernels and the GEMM variants at the matching row count
        rows = B * T
        x = torch.randn(rows, D, device=dev)
        g, bta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
        y, mean, rstd = ops.layernorm_fwd(x, g, bta)
        dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
        dx = ops.layernorm_bwd(torch.randn_like(x), x, g, mean, rstd, dg, db)
        cs = torch.zeros(D, device=dev)
        yb = ops.cast_bf16_colsum(x, cs)
        w = torch.randn(D, D, device=dev).bfloat16()
        c = ops.gemm(yb, w, bias=torch.zeros(D, device=dev), relu=True, out_dtype=torch.bfloat16)            # forward
        dw = torch.zeros(D, D, device=dev)
        ops.gemm(c, yb, a_mn_major=True, b_mn_major=True, out=dw, accumulate=True, split_k=2 if rows >= 512 else 1)      # wgrad, split-K atomics
        dxg = ops.gemm(c, w, b_mn_major=True, resid=x)                                                          # dgrad + residual
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all() and torch.isfinite(dqk.float()).all() and torch.isfinite(dxg).all()
    print(f"ok B={B} T={T} H={H} R={R} bucket={bucket} causal={causal} pad={pad}")
