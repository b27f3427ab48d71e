"""Debug: clock64 stamps of CTA (0,0,0) of the cross-attention forward kernel.
Build: RTTS_LIB_NAME=libreformer_b200_trace.so RTTS_DEFS=-DRTTS_TRACE python reformer_tts_b200/csrc/build.py
Run:   RTTS_LIB=$PWD/reformer_tts_b200/libreformer_b200_trace.so python tools/trace_xattn.py"""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reformer_tts_b200 import ops, _lib
lib = _lib.load()
B, T, S, H = 20, 1024, 256, 8; D = H * 64
q = torch.randn(B, T, D, device="cuda").bfloat16(); kv = torch.randn(B, S, 2 * D, device="cuda").bfloat16()
keep = torch.ones(B, S, dtype=torch.uint8, device="cuda"); keep[:, 200:] = 0
seed = torch.tensor([12345], dtype=torch.int64, device="cuda")
for _ in range(3): ops.xattn_fwd(q, kv[..., :D], kv[..., D:], keep, H, 0.125, 0.15, seed)
tr = torch.zeros(16, dtype=torch.int64, device="cuda")
lib.rtts_debug_set_xattn_trace.argtypes = [ctypes.c_void_p]
lib.rtts_debug_set_xattn_trace(ctypes.c_void_p(tr.data_ptr()))
ops.xattn_fwd(q, kv[..., :D], kv[..., D:], keep, H, 0.125, 0.15, seed); torch.cuda.synchronize()
lib.rtts_debug_set_xattn_trace(None)
t = tr.cpu().tolist()
names = ["start", "copies issued", "loads landed + tmem", "S done", "pass 1 (max) done", "pass 2 (exp, P) done", "PV done", "stored"]
for n, v in zip(names, t): print(f"{n:24s} {v - t[0]}")
