"""Per-shape timing of the HBM-bound row-wise kernels (CUDA events, 50 launches each, rotating buffers larger than L2):
python tools/time_rowwise.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reformer_tts_b200 import ops  # noqa: E402
dev = "cuda"
torch.manual_seed(0)


def timeit(fn, n=50):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n


for rows, cols in [(20480, 512), (20480, 1024), (5120, 512)]:
    nbuf = max(2, int(400e6 // (rows * cols * 4)))
    xs = [torch.randn(rows, cols, device=dev) for _ in range(nbuf)]
    cs = torch.zeros(cols, device=dev)
    g, bta = torch.ones(cols, device=dev), torch.zeros(cols, device=dev)
    us = timeit(lambda i: ops.cast_bf16_colsum(xs[i % nbuf], cs))
    mb = rows * cols * 6 / 1e6
    print(f"cast_bf16_colsum [{rows}x{cols}]  {us:7.1f} us  {mb / us / 1e3 * 1e3:7.0f} GB/s")
    us = timeit(lambda i: ops.layernorm_fwd(xs[i % nbuf], g, bta))
    print(f"layernorm_fwd    [{rows}x{cols}]  {us:7.1f} us  {mb / us / 1e3 * 1e3:7.0f} GB/s")
    y, mean, rstd = ops.layernorm_fwd(xs[0], g, bta)
    dg, db = torch.zeros(cols, device=dev), torch.zeros(cols, device=dev)
    us = timeit(lambda i: ops.layernorm_bwd(xs[(i + 1) % nbuf], xs[i % nbuf], g, mean, rstd, dg, db))
    print(f"layernorm_bwd    [{rows}x{cols}]  {us:7.1f} us  {rows * cols * 12 / 1e6 / us / 1e3 * 1e3:7.0f} GB/s")
