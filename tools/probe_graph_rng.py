"""Probe: graph-safe RNG snapshot / replay inside a captured CUDA graph (what Deterministic needs)."""
import torch
dev = torch.device("cuda", 0)
torch.cuda.init()
gen = torch.cuda.default_generators[0]
torch.manual_seed(0)
a = torch.empty(8, device=dev); b = torch.empty(8, device=dev); c = torch.empty(8, device=dev); d = torch.empty(8, device=dev)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3):
        a.copy_(torch.randn(8, device=dev))
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    a.copy_(torch.randn(8, device=dev))
    snap = gen.graphsafe_get_state().clone_state()
    b.copy_(torch.randn(8, device=dev))
    cur = gen.graphsafe_get_state()
    gen.graphsafe_set_state(snap)
    c.copy_(torch.randn(8, device=dev))
    gen.graphsafe_set_state(cur)
    d.copy_(torch.randn(8, device=dev))
for it in range(3):
    g.replay()
    torch.cuda.synchronize()
    print(it, "c==b (replayed draw)", torch.equal(c, b), "a", a[:3].tolist(), "d!=b", not torch.equal(d, b), "d!=a", not torch.equal(d, a))
