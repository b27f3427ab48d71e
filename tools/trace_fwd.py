"""Debug: per-role clock64 timeline of CTA 0 of the forward attention kernel."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reformer_tts_b200 import ops, _lib
lib = _lib.load()
B, T, H, R, bucket = 20, 1024, 8, 8, 64
dev = "cuda"
torch.manual_seed(0)
qkv = torch.randn(B, T, 2 * H * 64, device=dev).bfloat16()
qk, v = qkv[..., :H * 64], qkv[..., H * 64:]
nb = T // bucket
rot = torch.randn(1, 64, R, nb // 2, device=dev)
spec = ops.LSHSpec.reformer_pytorch(64, True)
buckets, sumsq = ops.lsh_hash(qk, rot, H, R, nb, return_sumsq=True)
sticker, undo = ops.lsh_sort(buckets, T, R, nb)
for _ in range(3):
    ops.lsh_attn_fwd(qk, v, sticker, None, spec, H, R, bucket, sumsq=sumsq)
trace = torch.zeros(3 * 32 * 8, dtype=torch.int64, device=dev)
lib.rtts_debug_set_fwd_trace.argtypes = [ctypes.c_void_p]
lib.rtts_debug_set_fwd_trace(ctypes.c_void_p(trace.data_ptr()))
ops.lsh_attn_fwd(qk, v, sticker, None, spec, H, R, bucket, sumsq=sumsq)
torch.cuda.synchronize()
lib.rtts_debug_set_fwd_trace(None)
t = trace.cpu().view(3, 32, 8)
t0 = int(t[t > 0].min())
names = {0: ["kv_full ok", "S issued", "p_full ok", "PV issued"], 1: ["start", "stk loaded", "kv_free ok", "arrived"], 2: ["start", "s_full ok", "P done", "o_full ok", "epi done", "fast done"]}
for n in range(4, 14):
    print(f"--- tile {n}")
    for role, rn in ((1, "loader"), (0, "mma"), (2, "softmax")):
        print(f"  {rn:8s}", "  ".join(f"{nm}={(int(t[role, n, k]) - t0) if nm != "slowmask" else hex(int(t[role, n, k]))}" for k, nm in enumerate(names[role])))
