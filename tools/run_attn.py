"""Launches the LSH kernels alone at a reference-config shape (for ncu): python tools/run_attn.py [B T H R bucket causal iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reformer_tts_b200 import ops  # noqa: E402

a = [int(x) for x in sys.argv[1:]]
B, T, H, R, bucket, causal, iters, pad = (a + [20, 1024, 8, 8, 64, 1, 3, 0][len(a):])[:8]
dev = "cuda"
torch.manual_seed(0)
qkv = torch.randn(B, T, 2 * H * 64, device=dev).bfloat16()
qk, v = qkv[..., :H * 64], qkv[..., H * 64:]
dout = torch.randn(B, T, H * 64, device=dev).bfloat16()
nb = T // bucket
rot = torch.randn(1, 64, R, nb // 2, device=dev)
spec = ops.LSHSpec.reformer_pytorch(64, bool(causal))
mask = None
if pad:      # the last `pad` positions of every sequence are padding (uint8 mask, 1 = valid)
    mask = torch.ones(B, T, dtype=torch.uint8, device=dev)
    mask[:, T - pad:] = 0
for it in range(iters):
    buckets = ops.lsh_hash(qk, rot, H, R, nb)
    sticker, undo = ops.lsh_sort(buckets, T, R, nb)
    o, lse_r = ops.lsh_attn_fwd(qk, v, sticker, mask, spec, H, R, bucket)
    out, lse = ops.lsh_merge_fwd(o, lse_r)
    delta = ops.lsh_delta(dout, out, H)
    dqk, dv = ops.lsh_attn_bwd(qk, v, sticker, undo, mask, spec, dout, lse, delta, H, R, bucket)
torch.cuda.synchronize()
timer = ops.KernelTimer()
ops.set_kernel_timer(timer)
for it in range(5):
    buckets = ops.lsh_hash(qk, rot, H, R, nb)
    sticker, undo = ops.lsh_sort(buckets, T, R, nb)
    o, lse_r = ops.lsh_attn_fwd(qk, v, sticker, mask, spec, H, R, bucket)
    out, lse = ops.lsh_merge_fwd(o, lse_r)
    delta = ops.lsh_delta(dout, out, H)
    dqk, dv = ops.lsh_attn_bwd(qk, v, sticker, undo, mask, spec, dout, lse, delta, H, R, bucket)
torch.cuda.synchronize()
flops = 8.0 * B * R * T * bucket * H * 64
for k, v_ in timer.summary().items():
    extra = ""
    if k.startswith("lsh_attn_fwd"):
        extra = f"  {flops / v_['avg_ms'] / 1e9:.1f} TFLOP/s"
    if k.startswith("lsh_attn_bwd"):
        extra = f"  {2.5 * flops / v_['avg_ms'] / 1e9:.1f} TFLOP/s"
    print(f"{k:32s} avg {v_['avg_ms'] * 1e3:9.1f} us{extra}")
