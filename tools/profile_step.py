"""Per-kernel breakdown of one training step (diagnostic): CUDA-event timing of every libreformer_b200 launch and a
torch.profiler table of all kernels.  Usage: python tools/profile_step.py [config] [batch]"""
import os
import sys
import json

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from reformer_tts_b200 import ops  # noqa: E402
from reformer_tts_b200.model import ReformerTTS, config as C  # noqa: E402
from reformer_tts_b200.model.loss import TTSLoss  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_CONFIG
batch = int(sys.argv[2]) if len(sys.argv) > 2 else C.REFERENCE_BATCH[name]
dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
torch.manual_seed(42)
model = ReformerTTS(**C.reference_model_kwargs(name)).to(dev).train()
loss_fn = TTSLoss(torch.tensor(5.)).to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)
data = {k: v.to(dev) for k, v in bench.synthetic_batch(batch).items()}
for _ in range(3):
    bench.train_step(model, loss_fn, opt, data)
torch.cuda.synchronize()

timer = ops.KernelTimer()
ops.set_kernel_timer(timer)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
bench.train_step(model, loss_fn, opt, data)
b.record()
torch.cuda.synchronize()
ops.set_kernel_timer(None)
step_ms = a.elapsed_time(b)
summ = timer.summary()
tot = sum(v["total_ms"] for v in summ.values())
print(f"step {step_ms:.2f} ms (with event overhead); libreformer_b200 kernels {tot:.2f} ms in {sum(v['count'] for v in summ.values())} launches")
for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["total_ms"]):
    print(f"  {k:48s} n={v['count']:4d} total={v['total_ms']:8.3f} ms avg={v['avg_ms'] * 1e3:9.1f} us")

from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    bench.train_step(model, loss_fn, opt, data)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))

# kernels only, by name
from torch.autograd import DeviceType  # noqa: E402
ks = [e for e in prof.key_averages() if e.device_type == DeviceType.CUDA]
tot_k = sum(e.self_device_time_total for e in ks)
print(f"\nkernels only: {tot_k / 1e3:.2f} ms in {sum(e.count for e in ks)} launches")
for e in sorted(ks, key=lambda e: -e.self_device_time_total)[:45]:
    print(f"  {e.self_device_time_total / 1e3:8.3f} ms {100 * e.self_device_time_total / tot_k:5.1f}%  n={e.count:4d}  {e.key[:110]}")

# where the non-library kernels come from: kernels grouped by (kernel name, innermost frame inside this repository)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof2:
    bench.train_step(model, loss_fn, opt, data)
    torch.cuda.synchronize()
from collections import defaultdict  # noqa: E402
agg = defaultdict(lambda: [0, 0.0])
evs = prof2.events()
for e in evs:
    if e.device_type == DeviceType.CUDA or not e.kernels:
        continue
    frame = next((f for f in (e.stack or []) if "/repo/" in f and "profile_step" not in f), "(autograd / optimizer)")
    for k in e.kernels:
        if "rtts" in k.name or "spin_kernel" in k.name:
            continue
        key = (k.name[:60], frame.split("/repo/")[-1][:80])
        agg[key][0] += 1
        agg[key][1] += k.duration
print("\nnon-library kernels by call site:")
for (kn, fr), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"  {t / 1e3:7.3f} ms n={n:4d}  {kn:60s} {fr}")
