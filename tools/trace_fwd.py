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
trace = torch.zeros(4 * 32 * 8 + 148 * 32, dtype=torch.int64, device=dev)
lib.rtts_debug_set_fwd_trace.argtypes = [ctypes.c_void_p]
lib.rtts_debug_set_fwd_trace(ctypes.c_void_p(trace.data_ptr()))
ops.lsh_attn_fwd(qk, v, sticker, None, spec, H, R, bucket, sumsq=sumsq)
torch.cuda.synchronize()
lib.rtts_debug_set_fwd_trace(None)
per_cta = trace.cpu()[4 * 32 * 8:].view(148, 32)
t = trace.cpu()[:4 * 32 * 8].view(4, 32, 8)
t0 = int(t[t > 0].min())
names = {0: ["kv_full ok", "S issued", "p_full ok", "PV issued", "loop top", "o_free ok", "PV mmas out"], 1: ["start", "stk loaded", "kv_free ok", "arrived"], 2: ["start", "s_full ok", "P done", "full ok", "S loaded", "chunks done", "fin stored", "P stored"], 3: ["o_full ok", "stored"]}
for n in range(0, 16):
    print(f"--- tile {n}")
    for role, rn in ((1, "loader"), (0, "mma"), (2, "softmax"), (3, "epilogue")):
        print(f"  {rn:8s}", "  ".join(f"{nm}={(int(t[role, n, k]) - t0) if nm not in ("-",) else int(t[role, n, k])}" for k, nm in enumerate(names[role])))

print("per-CTA cycles until role done: softmax(w0) softmax(w8) epilogue(w16) loader(w20) mma(w24)")
for cta in list(range(0, 148, 12)) + [147]:
    print(cta, [int(per_cta[cta, w]) for w in (0, 8, 16, 20, 24)])
print("loader done: min/median/max", int(per_cta[:, 20].min()), int(per_cta[:, 20].median()), int(per_cta[:, 20].max()))
print("mma done: min/median/max", int(per_cta[:, 24].min()), int(per_cta[:, 24].median()), int(per_cta[:, 24].max()))
