"""BASELINE.json configs[4]: LSH-attention long-sequence sweep, 1k-16k positions, 8 heads, 4 hash rounds, bucket 64.
Times the LSH core (hash -> sort -> chunked attention -> merge, and its backward) and the whole LSHSelfAttention layer
(LayerNorm + projections + core + output projection, forward + backward) on ONE GPU; (batch, head) units are independent, so
multi-GPU scaling of this workload is plain replication: under torchrun every rank runs its own ~16k positions on its own
GPU, no data-path collective; times are the maximum over ranks (bracketed by barriers) and rates are the sum over ranks.
Prints one JSON line per length (rank 0).  python tools/sweep_lsh.py | torchrun --nproc-per-node N tools/sweep_lsh.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reformer_tts_b200 import ops  # noqa: E402
from reformer_tts_b200.lsh_attention import LSHSelfAttention  # noqa: E402

H, R, bucket, D = 8, 4, 64, 512
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = "cuda"
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
peak = 1396.7
if os.path.exists("MEASURED_PEAKS.json"):
    peak = json.load(open("MEASURED_PEAKS.json"))["bf16_tflops_sustained"]


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    if world > 1:      # slowest rank
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    return ms


for T in (1024, 2048, 4096, 8192, 16384):
    B = max(1, 16384 // T)          # keep ~16k positions in flight so every length fills the GPU
    torch.manual_seed(rank)
    qkv = torch.randn(B, T, 2 * D, device=dev).bfloat16()
    qk, v = qkv[..., :D], qkv[..., D:]
    dout = torch.randn(B, T, D, device=dev).bfloat16()
    nb = T // bucket
    rot = torch.randn(1, 64, R, nb // 2, device=dev)
    spec = ops.LSHSpec.reformer_pytorch(64, True)
    state = {}

    def core_fwd():
        buckets, sumsq = ops.lsh_hash(qk, rot, H, R, nb, return_sumsq=True)
        sticker, undo = ops.lsh_sort(buckets, T, R, nb)
        o, lse_r = ops.lsh_attn_fwd(qk, v, sticker, None, spec, H, R, bucket, sumsq=sumsq)
        out, lse = ops.lsh_merge_fwd(o, lse_r)
        state.update(sticker=sticker, undo=undo, sumsq=sumsq, out=out, lse=lse)

    def core_bwd():
        delta = ops.lsh_delta(dout, state["out"], H)
        ops.lsh_attn_bwd(qk, v, state["sticker"], state["undo"], None, spec, dout, state["lse"], delta, H, R, bucket, sumsq=state["sumsq"])

    def attn_only():
        ops.lsh_attn_fwd(qk, v, state["sticker"], None, spec, H, R, bucket, sumsq=state["sumsq"])

    ms_f = timed(core_fwd)
    ms_b = timed(core_bwd)
    ms_a = timed(attn_only)
    layer = LSHSelfAttention(D, heads=H, bucket_size=bucket, n_hashes=R, causal=True).to(dev)
    norm = torch.nn.LayerNorm(D).to(dev)
    x = torch.randn(B, T, D, device=dev, requires_grad=True)
    dy = torch.randn(B, T, D, device=dev)

    def layer_step():
        x.grad = None
        layer(x, norm=norm).backward(dy)

    ms_l = timed(layer_step, iters=5)
    flops = 8.0 * B * R * T * bucket * D
    if rank == 0:
      print(json.dumps({"n_gpus": world, "T": T, "batch_per_gpu": B, "heads": H, "rounds": R, "bucket": bucket, "core_fwd_ms": round(ms_f, 4), "core_bwd_ms": round(ms_b, 4),
                      "attn_kernel_ms": round(ms_a, 4), "attn_kernel_tflops": round(flops / ms_a / 1e9, 1), "attn_kernel_frac_of_peak": round(flops / ms_a / 1e9 / peak, 4),
                      "layer_fwd_bwd_ms": round(ms_l, 4), "positions_per_s_layer_fwd_bwd": round(world * B * T / ms_l * 1e3),
                      "positions_per_s_core_fwd_bwd": round(world * B * T / (ms_f + ms_b) * 1e3)}), flush=True)
if world > 1:
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)
