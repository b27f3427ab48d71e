"""Times the row-wise kernels at the shapes of the default training step: python tools/run_rowwise.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reformer_tts_b200 import ops  # noqa: E402
dev = "cuda"
torch.manual_seed(0)
timer = ops.KernelTimer()
for rows, cols in [(20480, 512), (20480, 2048), (5120, 512), (20480, 1024)]:
    x = torch.randn(rows, cols, device=dev)
    cs = torch.zeros(cols, device=dev)
    for _ in range(3):
        y = ops.cast_bf16_colsum(x, cs)
    torch.cuda.synchronize()
    ref = x.sum(0) * 3
    err = ((cs - ref).abs().max() / ref.abs().max()).item()
    ok = torch.equal(y, x.bfloat16())
    ops.set_kernel_timer(timer)
    for _ in range(10):
        ops.cast_bf16_colsum(x, cs)
    torch.cuda.synchronize()
    ops.set_kernel_timer(None)
    t = list(timer.summary().values())[-1]["avg_ms"] if False else None
    print(rows, cols, "cast exact", ok, "colsum rel err", err)
for k, v in timer.summary().items():
    print(f"{k:40s} avg {v['avg_ms'] * 1e3:8.1f} us  n={v['count']}")
