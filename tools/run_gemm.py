"""Times rtts_gemm_bf16 at the shapes of the default training step: python tools/run_gemm.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reformer_tts_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
M = 20480
cases = [("fwd  x[M,512]  . W[2048,512]^T", M, 2048, 512, False, False, 1), ("fwd  h[M,2048] . W[512,2048]^T", M, 512, 2048, False, False, 1),
         ("fwd  x[M,512]  . Wqkv[1024,512]^T", M, 1024, 512, False, False, 1), ("fwd  o[M,512]  . Wo[512,512]^T", M, 512, 512, False, False, 1),
         ("dgrad dy[M,2048] . W[2048,512] (B MN-major)", M, 512, 2048, False, True, 1), ("dgrad dy[M,512] . W[512,2048] (B MN-major)", M, 2048, 512, False, True, 1),
         ("wgrad dY^T X: [2048,M]x[M,512] (both MN-major)", 2048, 512, M, True, True, 0), ("wgrad dY^T X: [512,M]x[M,2048]", 512, 2048, M, True, True, 0),
         ("wgrad dY^T X: [1024,M]x[M,512]", 1024, 512, M, True, True, 0)]
timer = ops.KernelTimer()
for name, m, n, k, amn, bmn, split in cases:
    a = torch.randn((k, m) if amn else (m, k), device=dev).bfloat16()
    b = torch.randn((k, n) if bmn else (n, k), device=dev).bfloat16()
    kw = {}
    if split == 0:      # let the library pick split-K as the training step does (atomic fp32 accumulate)
        kw = dict(out=torch.zeros(m, n, device=dev), accumulate=True, split_k=max(1, min(16, (2 * 148 * 128 * 128) // (m * n))))
        while k % (64 * kw["split_k"]):
            kw["split_k"] -= 1
    for _ in range(3):
        ops.gemm(a, b, a_mn_major=amn, b_mn_major=bmn, **kw)
    torch.cuda.synchronize()
    ops.set_kernel_timer(timer)
    n0 = len(timer.summary())
    for _ in range(10):
        ops.gemm(a, b, a_mn_major=amn, b_mn_major=bmn, **kw)
    torch.cuda.synchronize()
    ops.set_kernel_timer(None)
for (tag, v), case in zip(timer.summary().items(), cases):
    name, m, n, k = case[:4]
    print(f"{name:52s} {tag:34s} avg {v['avg_ms'] * 1e3:8.1f} us  {2.0 * m * n * k / v['avg_ms'] / 1e9:7.1f} TFLOP/s")
