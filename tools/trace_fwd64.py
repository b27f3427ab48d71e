"""Debug: per-role clock64 timeline of CTA 0 of the block-streaming forward attention kernel (bucket 64).
Build: RTTS_LIB_NAME=libreformer_b200_trace.so RTTS_DEFS=-DRTTS_TRACE python reformer_tts_b200/csrc/build.py
Run:   RTTS_LIB=$PWD/reformer_tts_b200/libreformer_b200_trace.so python tools/trace_fwd64.py"""
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
names = {1: ["start", "slot free", "copies issued", "meta done"], 0: ["S: ready", "S: committed", "PV: p_full", "PV: d free", "PV: committed"],
         2: ["loop top", "s_full", "w0 arrived", "w3 arrived", "ld01 in", "c01 done", "ld23 in", "c23 done"], 3: ["lb: pv_done", "lb: handed", "main: pv_done", "main: x_full", "main: stage free", "main: stored", "main: tmem in", "main: rows staged"]}
for n in range(int(sys.argv[1]) if len(sys.argv) > 1 else 24):
    print(f"--- entry {n}")
    for role, rn in ((1, "loader"), (0, "mma"), (2, "softmax"), (3, "epilogue")):
        print(f"  {rn:8s}", "  ".join(f"{nm}={int(t[role, n, k]) - t0 if int(t[role, n, k]) else '-'}" for k, nm in enumerate(names[role])))
