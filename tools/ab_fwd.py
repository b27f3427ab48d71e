"""A/B of the three bucket-64 forward kernels (paired-chunk = default, 128-query tiles, block-streaming) at the default workload's shapes."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reformer_tts_b200 import ops, _lib
lib = _lib.load()
lib.rtts_debug_set_fwd_kernel.argtypes = [ctypes.c_int]
dev = "cuda"
for (B, T, H, R, causal, pad) in [(20, 1024, 8, 8, 1, 0), (20, 256, 8, 8, 0, 40), (16, 1024, 8, 4, 1, 0), (8, 2048, 8, 4, 1, 0), (16, 1024, 8, 4, 0, 0)]:
    torch.manual_seed(0)
    qkv = torch.randn(B, T, 2 * H * 64, device=dev).bfloat16()
    qk, v = qkv[..., :H * 64], qkv[..., H * 64:]
    nb = T // 64
    rot = torch.randn(1, 64, R, nb // 2, device=dev)
    spec = ops.LSHSpec.reformer_pytorch(64, bool(causal))
    mask = None
    if pad:
        mask = torch.ones(B, T, dtype=torch.uint8, device=dev)
        mask[:, T - pad:] = 0
    buckets, sumsq = ops.lsh_hash(qk, rot, H, R, nb, return_sumsq=True)
    sticker, undo = ops.lsh_sort(buckets, T, R, nb)
    outs = []
    for which in (0, 1, 2):
        lib.rtts_debug_set_fwd_kernel(which)
        for _ in range(3):
            o, lse = ops.lsh_attn_fwd(qk, v, sticker, mask, spec, H, R, 64, sumsq=sumsq)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(20):
            o, lse = ops.lsh_attn_fwd(qk, v, sticker, mask, spec, H, R, 64, sumsq=sumsq)
        ev[1].record()
        torch.cuda.synchronize()
        us = ev[0].elapsed_time(ev[1]) * 1e3 / 20
        fl = 8.0 * B * R * T * 64 * H * 64
        outs.append((o.float(), lse))
        print(f"B={B} T={T} R={R} causal={causal} pad={pad} kernel={('pair', 'tile', 'block')[which]}: {us:7.1f} us  {fl / us / 1e6:6.1f} TFLOP/s")
    lib.rtts_debug_set_fwd_kernel(0)
    for w in (1, 2):
        print(f"   pair vs {('pair', 'tile', 'block')[w]}: max |o diff|", (outs[0][0] - outs[w][0]).abs().max().item(), " max |lse diff|", (outs[0][1] - outs[w][1]).abs().max().item())
