"""Debug: clock64 timeline of CTA 0 of the backward attention kernel."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reformer_tts_b200 import ops, _lib
lib = _lib.load()
B, T, H, R, bucket = 20, 1024, 8, 8, 64
dev = "cuda"
torch.manual_seed(0)
qkv = torch.randn(B, T, 2 * H * 64, device=dev).bfloat16()
qk, v = qkv[..., :H * 64], qkv[..., H * 64:]
dout = torch.randn(B, T, H * 64, device=dev).bfloat16()
nb = T // bucket
rot = torch.randn(1, 64, R, nb // 2, device=dev)
spec = ops.LSHSpec.reformer_pytorch(64, True)
buckets, sumsq = ops.lsh_hash(qk, rot, H, R, nb, return_sumsq=True)
sticker, undo = ops.lsh_sort(buckets, T, R, nb)
o, lse_r = ops.lsh_attn_fwd(qk, v, sticker, None, spec, H, R, bucket, sumsq=sumsq)
out, lse = ops.lsh_merge_fwd(o, lse_r)
delta = ops.lsh_delta(dout, out, H)
for _ in range(2):
    ops.lsh_attn_bwd(qk, v, sticker, undo, None, spec, dout, lse, delta, H, R, bucket, sumsq=sumsq)
trace = torch.zeros(32, dtype=torch.int64, device=dev)
lib.rtts_debug_set_bwd_trace.argtypes = [ctypes.c_void_p]
lib.rtts_debug_set_bwd_trace(ctypes.c_void_p(trace.data_ptr()))
ops.lsh_attn_bwd(qk, v, sticker, undo, None, spec, dout, lse, delta, H, R, bucket, sumsq=sumsq)
torch.cuda.synchronize()
lib.rtts_debug_set_bwd_trace(None)
t = trace.cpu().tolist()
names = {0: "start", 1: "cp.async issued", 2: "meta written", 3: "gather synced", 4: "qb0 scores ready", 5: "qb0 elementwise done", 6: "qb0 mma issued",
         7: "qb1 scores ready", 8: "qb1 elementwise done", 9: "qb1 mma issued", 10: "qb2 scores ready", 11: "qb2 elementwise done", 12: "qb2 mma issued",
         20: "accumulators ready", 21: "epilogue stores issued", 22: "end"}
prev = t[0]
for k in sorted(names):
    if t[k]:
        print(f"{names[k]:26s} +{t[k] - prev:6d}   (t={t[k] - t[0]})")
        prev = t[k]
