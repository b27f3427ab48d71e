"""Debug: per-role clock64 timeline of CTA 0 of the paired-chunk forward attention kernel (bucket 64).
Build: RTTS_LIB_NAME=libreformer_b200_trace.so RTTS_DEFS=-DRTTS_TRACE python reformer_tts_b200/csrc/build.py
Run:   RTTS_LIB=$PWD/reformer_tts_b200/libreformer_b200_trace.so python tools/trace_fwd64p.py [tiles]"""
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
trace = torch.zeros(4 * 64 * 8, dtype=torch.int64, device=dev)
lib.rtts_debug_set_fwd_trace.argtypes = [ctypes.c_void_p]
lib.rtts_debug_set_fwd_trace(ctypes.c_void_p(trace.data_ptr()))
ops.lsh_attn_fwd(qk, v, sticker, None, spec, H, R, bucket, sumsq=sumsq)
torch.cuda.synchronize()
lib.rtts_debug_set_fwd_trace(None)
t = trace.cpu().view(4, 64, 8)
t0 = int(t[t > 0].min())
names = {1: ["K start", "K slot free", "K copies issued", "K meta done", "V start", "V slot free", "V copies issued", "V done"],
         0: ["S: top", "S: landed", "S: region free", "S: committed", "PV: top", "PV: p_full", "PV: committed"],
         3: ["meta read", "ld0", "blk0 done", "ld1", "blk1 done", "st_wait done"],
         2: ["top", "s_full", "w0 p-arrived", "w3 p-arrived", "o_full", "o_free arrived", "stored"]}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for i in range(2 * n + 2):
    print(f"entry {i:2d} loader ", "  ".join(f"{nm}={int(t[1, i, k]) - t0 if int(t[1, i, k]) else '-'}" for k, nm in enumerate(names[1])))
for i in range(n):
    print(f"--- tile {i}")
    for role, rn in ((0, "mma"), (2, "softmax"), (3, "w0 detail")):
        print(f"  {rn:8s}", "  ".join(f"{nm}={int(t[role, i, k]) - t0 if int(t[role, i, k]) else '-'}" for k, nm in enumerate(names[role])))
