"""GPU bring-up probe: runs every kernel against a torch/oracle reference, each group in its own
subprocess (a device trap poisons the CUDA context), and prints error statistics.  Not a test: a
diagnostic for the first runs on real hardware.  Usage: python tools/probe_kernels.py [group ...]"""
import subprocess
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
GROUPS = ["gemm_tn", "gemm_mn", "gemm_epi", "rowwise", "hash", "attn64", "attn128", "bwd64", "bwd128"]


def rel(a, b):
    import torch
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def run(group):
    import torch
    from reformer_tts_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    if group.startswith("gemm"):
        for (m, n, k) in [(128, 128, 64), (256, 128, 512), (1024, 2048, 512), (512, 512, 2048)]:
            a = torch.randn(m, k, device=dev).bfloat16()
            b = torch.randn(n, k, device=dev).bfloat16()
            ref = a.float() @ b.float().t()
            if group == "gemm_tn":
                c = ops.gemm(a, b)
                print(group, (m, n, k), "rel", rel(c, ref), flush=True)
            elif group == "gemm_mn":
                at, bt = a.t().contiguous(), b.t().contiguous()
                for amn, bmn in [(True, False), (False, True), (True, True)]:
                    c = ops.gemm(at if amn else a, bt if bmn else b, a_mn_major=amn, b_mn_major=bmn)
                    print(group, (m, n, k), (amn, bmn), "rel", rel(c, ref), flush=True)
                if k >= 256:
                    acc = torch.ones(m, n, device=dev)
                    ops.gemm(at, bt, a_mn_major=True, b_mn_major=True, out=acc, accumulate=True, split_k=4)
                    print(group, (m, n, k), "splitk4 atomic rel", rel(acc, ref + 1), flush=True)
            else:
                bias = torch.randn(n, device=dev)
                gate = torch.randn(m, n, device=dev).bfloat16()
                cs = torch.zeros(n, device=dev)
                c = ops.gemm(a, b, bias=bias, relu=True, out_dtype=torch.bfloat16)
                print(group, (m, n, k), "bias+relu bf16 rel", rel(c, torch.relu(ref + bias)), flush=True)
                c = ops.gemm(a, b, gate=gate, colsum=cs, out_dtype=torch.bfloat16)
                refg = ref * (gate.float() > 0)
                print(group, (m, n, k), "gate rel", rel(c, refg), "colsum rel", rel(cs, refg.sum(0)), flush=True)
    elif group == "rowwise":
        x = torch.randn(4096, 512, device=dev) * 2 + 0.5
        g, bta = torch.randn(512, device=dev), torch.randn(512, device=dev)
        y, mean, rstd = ops.layernorm_fwd(x, g, bta)
        xr = x.clone().requires_grad_(True); gr = g.clone().requires_grad_(True); br = bta.clone().requires_grad_(True)
        yr = torch.nn.functional.layer_norm(xr, (512,), gr, br, 1e-5)
        print("ln fwd rel", rel(y, yr), flush=True)
        dy = torch.randn_like(x)
        yr.backward(dy)
        dg, db = torch.zeros(512, device=dev), torch.zeros(512, device=dev)
        dx = ops.layernorm_bwd(dy, x, g, mean, rstd, dg, db)
        print("ln bwd dx", rel(dx, xr.grad), "dg", rel(dg, gr.grad), "db", rel(db, br.grad), flush=True)
        cs = torch.zeros(512, device=dev)
        yb = ops.cast_bf16_colsum(x, cs)
        print("cast", rel(yb, x.bfloat16()), "colsum", rel(cs, x.sum(0)), flush=True)
        a = torch.randn(3, 256, 512, device=dev).bfloat16(); o = torch.randn(3, 256, 512, device=dev).bfloat16()
        d = ops.lsh_delta(a, o, 8)
        dref = (a.float() * o.float()).view(3, 256, 8, 64).sum(-1).permute(0, 2, 1)
        print("delta rel", rel(d, dref), flush=True)
    elif group in ("hash", "sort"):
        from oracle import lsh_core
        B, T, H, R = 3, 512, 8, 4
        for nb, per_head, pad in [(8, False, False), (16, True, True), (256, False, False)]:
            qk = torch.randn(B, T, H * 64, device=dev).bfloat16()
            rot = torch.randn(H if per_head else 1, 64, R, nb // 2, device=dev)
            mask = None
            if pad:
                mask = torch.ones(B, T, dtype=torch.bool, device=dev); mask[0, -37:] = False
            bk = ops.lsh_hash(qk, rot, H, R, nb, None if mask is None else mask.to(torch.uint8), pad)
            q = qk.float().view(B, T, H, 64).transpose(1, 2).reshape(B * H, T, 64).cpu()
            rt = rot.cpu()
            rt = rt[None].expand(B, -1, -1, -1, -1).reshape(B * H, 64, R, nb // 2) if per_head else rt
            m = None if mask is None else mask[:, None, :].expand(B, H, T).reshape(B * H, T).cpu()
            ref = lsh_core.hash_buckets(q, rt, R, nb, m)
            mism = (bk.cpu().view(B * H, -1).long() != ref).sum().item()
            print("hash", nb, per_head, pad, "mismatches", mism, "of", ref.numel(), flush=True)
            ids = nb + 1 if pad else nb
            st, un = ops.lsh_sort(bk, T, R, ids)
            rs, ru = lsh_core.sort_buckets(bk.cpu().view(B * H, -1).long(), T)
            print("sort", nb, "sticker eq", bool((st.cpu().view(B * H, -1).long() == rs).all()),
                  "undo eq", bool((un.cpu().view(B * H, -1).long() == ru).all()), flush=True)
    elif group in ("attn64", "attn128", "merge"):
        from oracle import lsh_core
        bucket = 128 if group == "attn128" else 64
        B, T, H, R = 2, 512, 2, 4
        for impl in ("rp", "hf"):
            for causal in (False, True):
                for pad in (False, True):
                    qk = torch.randn(B, T, H * 64, device=dev).bfloat16()
                    v = torch.randn(B, T, H * 64, device=dev).bfloat16()
                    nb = T // bucket
                    buckets = torch.randint(0, nb, (B * H, R, T)) + nb * torch.arange(R).view(1, R, 1)
                    buckets = buckets.view(B * H, R * T)
                    sticker, undo = lsh_core.sort_buckets(buckets, T)
                    mask = None
                    if pad:
                        mask = torch.ones(B, T, dtype=torch.bool); mask[0, -100:] = False; mask[1, -3:] = False
                    ospec = (lsh_core.LSHSpec.reformer_pytorch if impl == "rp" else lsh_core.LSHSpec.huggingface)(64, causal)
                    gspec = (ops.LSHSpec.reformer_pytorch if impl == "rp" else ops.LSHSpec.huggingface)(64, causal)
                    q32 = qk.float().view(B, T, H, 64).transpose(1, 2).reshape(B * H, T, 64).cpu()
                    v32 = v.float().view(B, T, H, 64).transpose(1, 2).reshape(B * H, T, 64).cpu()
                    m_bh = None if mask is None else mask[:, None, :].expand(B, H, T).reshape(B * H, T)
                    so, slse = lsh_core.chunk_attention(q32, v32, sticker, bucket, R, ospec, m_bh)
                    out_ref, o_ref, lse_ref = lsh_core.unsort_and_merge(so, slse, undo, R)
                    o, lse = ops.lsh_attn_fwd(qk, v, sticker.to(torch.int32).to(dev).view(B, H, R * T),
                                              None if mask is None else mask.to(torch.uint8).to(dev), gspec, H, R, bucket)
                    torch.cuda.synchronize()
                    tag = (group, impl, "causal" if causal else "full", "pad" if pad else "nopad")
                    print(*tag, "o rel", rel(o.cpu().view(B * H, R, T, 64), o_ref),
                          "lse maxabs", (lse.cpu().view(B * H, R, T) - lse_ref).abs().max().item(), flush=True)
                    out, lse_tot = ops.lsh_merge_fwd(o, lse)
                    out_bh = out.float().view(B, T, H, 64).transpose(1, 2).reshape(B * H, T, 64).cpu()
                    print(*tag, "merged out rel", rel(out_bh, out_ref), "lse_tot maxabs",
                          (lse_tot.cpu().view(B * H, T) - torch.logsumexp(lse_ref, 1)).abs().max().item(), flush=True)
    elif group in ("bwd64", "bwd128"):
        from oracle import lsh_core
        bucket = 128 if group == "bwd128" else 64
        B, T, H, R = 2, 512, 2, 4
        for impl in ("rp", "hf"):
            for causal in (False, True):
                for pad in (False, True):
                    qk = torch.randn(B, T, H * 64, device=dev).bfloat16()
                    v = torch.randn(B, T, H * 64, device=dev).bfloat16()
                    dout = torch.randn(B, T, H * 64, device=dev).bfloat16()
                    nb = T // bucket
                    buckets = torch.randint(0, nb, (B * H, R, T)) + nb * torch.arange(R).view(1, R, 1)
                    buckets = buckets.view(B * H, R * T)
                    mask = None
                    if pad:
                        mask = torch.ones(B, T, dtype=torch.bool); mask[0, -100:] = False; mask[1, -3:] = False
                    ospec = (lsh_core.LSHSpec.reformer_pytorch if impl == "rp" else lsh_core.LSHSpec.huggingface)(64, causal)
                    gspec = (ops.LSHSpec.reformer_pytorch if impl == "rp" else ops.LSHSpec.huggingface)(64, causal)
                    to_bh = lambda x: x.float().view(B, T, H, 64).transpose(1, 2).reshape(B * H, T, 64).cpu()
                    q32 = to_bh(qk).requires_grad_(True); v32 = to_bh(v).requires_grad_(True)
                    m_bh = None if mask is None else mask[:, None, :].expand(B, H, T).reshape(B * H, T)
                    res = lsh_core.lsh_attention(q32, v32, buckets, bucket, R, ospec, m_bh)
                    (res["out"] * to_bh(dout)).sum().backward()
                    bk = buckets.to(torch.int32).to(dev).view(B, H, R * T)
                    sticker, undo = ops.lsh_sort(bk, T, R, nb)
                    mk = None if mask is None else mask.to(torch.uint8).to(dev)
                    o, lse_r = ops.lsh_attn_fwd(qk, v, sticker, mk, gspec, H, R, bucket)
                    out, lse = ops.lsh_merge_fwd(o, lse_r)
                    delta = ops.lsh_delta(dout, out, H)
                    dqk, dv = ops.lsh_attn_bwd(qk, v, sticker, undo, mk, gspec, dout, lse, delta, H, R, bucket)
                    torch.cuda.synchronize()
                    tag = (group, impl, "causal" if causal else "full", "pad" if pad else "nopad")
                    print(*tag, "dqk rel", rel(to_bh(dqk), q32.grad), "dv rel", rel(to_bh(dv), v32.grad), flush=True)
    import torch
    torch.cuda.synchronize()


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run(sys.argv[2])
        sys.exit(0)
    groups = sys.argv[1:] or GROUPS
    for g in groups:
        print(f"===== {g}", flush=True)
        try:
            r = subprocess.run([sys.executable, __file__, "--one", g], timeout=240, capture_output=True, text=True)
            print(r.stdout[-6000:], flush=True)
            if r.returncode != 0:
                print(f"[{g}] EXIT {r.returncode}\n{r.stderr[-3000:]}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"[{g}] TIMEOUT", flush=True)
