"""Stage-wise parity of every kernel against the CPU oracle, called through the C ABI (reformer_tts_b200.ops is a
ctypes shim over libreformer_b200.so).  Each stage is fed the oracle's inputs: integer stages must be
bit-exact, floating-point stages within the tolerances of tests/_util.py."""
import pytest
import torch

from _util import TOL, TOL_BF16_GRAD, TOL_BF16_STORED, TOL_FP32, bf16r, rel_l2, report, to_bh

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from reformer_tts_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def core():
    from oracle import lsh_core
    return lsh_core


@pytest.fixture(scope="module")
def rnd():
    from oracle import lsh_rounded
    return lsh_rounded


# ---------------------------------------------------------------------------------------------- hash / sort
@pytest.mark.parametrize("nb,per_head,pad,T", [(4, False, False, 256), (8, False, False, 1024), (16, False, False, 1024),
                                               (16, True, True, 1024), (8, True, False, 256), (256, False, False, 2048),
                                               (64, False, False, 4096), (128, True, False, 8192), (6, False, False, 384)])
def test_hash_bit_exact(ops, core, nb, per_head, pad, T):
    torch.manual_seed(nb + T)
    # projections per token P = R * nb / 2: 16 .. 256 are one tensor-pipe launch (rtts_lsh_hash_tc), 512 (nb = 256) two launches of two rounds
    B, H, R = 3 if T <= 2048 else 1, 8, 8 if nb < 64 else 4
    qk = torch.randn(B, T, H * 64, device=DEV).bfloat16()
    rot = torch.randn(H if per_head else 1, 64, R, nb // 2, device=DEV)
    mask = None
    if pad:
        mask = torch.ones(B, T, dtype=torch.bool, device=DEV)
        mask[0, -37:] = False
    got = ops.lsh_hash(qk, rot, H, R, nb, None if mask is None else mask.to(torch.uint8), pad).cpu().view(B * H, -1).long()
    q = to_bh(qk, H)
    rt = rot.cpu()
    rt = rt[None].expand(B, -1, -1, -1, -1).reshape(B * H, 64, R, nb // 2) if per_head else rt
    m = None if mask is None else mask[:, None, :].expand(B, H, T).reshape(B * H, T).cpu()
    want = core.hash_buckets(q, rt, R, nb, m)
    # exact wherever the fp64 top-2 margin is not a rounding-level tie (SURVEY.md 7 hard parts)
    proj = torch.einsum("ntd,ndri->nrti", q.double(), rt.double().expand(B * H, -1, -1, -1))
    top2 = torch.cat([proj, -proj], -1).topk(2, dim=-1).values
    safe = ((top2[..., 0] - top2[..., 1]) > 1e-5 * q.double().norm(dim=-1)[:, None, :]).reshape(B * H, -1)
    assert torch.equal(got[safe], want[safe])
    assert int((got != want).sum()) <= max(2, int(2e-5 * want.numel())), "more flips than fp32 ties can explain"


@pytest.mark.parametrize("T,R,nb,extra", [(256, 8, 4, 0), (1024, 8, 16, 0), (1024, 8, 16, 1), (2048, 4, 32, 0), (16384, 4, 256, 0), (128, 1, 2, 0)])
def test_sort_bit_exact(ops, core, T, R, nb, extra):
    torch.manual_seed(T + nb)
    rows = 5
    ids = nb + extra
    buckets = torch.randint(0, ids, (rows, R, T)) + ids * torch.arange(R).view(1, R, 1)
    buckets = buckets.view(rows, R * T)
    if T >= 1024:
        buckets[0] = (torch.arange(R).view(R, 1) * ids).expand(R, T).reshape(-1)      # everything in one bucket per round
    sticker, undo = ops.lsh_sort(buckets.to(torch.int32).to(DEV), T, R, ids)
    want_s, want_u = core.sort_buckets(buckets, T)
    assert torch.equal(sticker.cpu().long(), want_s)
    assert torch.equal(undo.cpu().long(), want_u)


# ---------------------------------------------------------------------------------------------- attention forward / merge / backward
def _attention_case(core, impl, causal, pad, bucket, B=2, T=512, H=2, R=4, seed=0, clustered=False):
    torch.manual_seed(seed)
    qk = torch.randn(B, T, H * 64, device=DEV).bfloat16()
    v = torch.randn(B, T, H * 64, device=DEV).bfloat16()
    nb = T // bucket
    buckets = torch.randint(0, nb, (B * H, R, T))
    if clustered:       # unbalanced buckets: chunks straddle bucket boundaries
        buckets = (buckets.float() ** 2 / nb).long().clamp_(0, nb - 1)
    buckets = (buckets + nb * torch.arange(R).view(1, R, 1)).view(B * H, R * T)
    mask = None
    if pad:
        mask = torch.ones(B, T, dtype=torch.bool)
        mask[0, -100:] = False
        mask[1, -3:] = False
    spec_o = (core.LSHSpec.reformer_pytorch if impl == "rp" else core.LSHSpec.huggingface)(64, causal)
    m_bh = None if mask is None else mask[:, None, :].expand(B, H, T).reshape(B * H, T)
    return dict(qk=qk, v=v, buckets=buckets, mask=mask, m_bh=m_bh, spec_o=spec_o, nb=nb, B=B, T=T, H=H, R=R, bucket=bucket, impl=impl, causal=causal)


def _gpu_spec(ops, c):
    return (ops.LSHSpec.reformer_pytorch if c["impl"] == "rp" else ops.LSHSpec.huggingface)(64, c["causal"])


@pytest.mark.parametrize("bucket", [64, 128])
@pytest.mark.parametrize("impl", ["rp", "hf"])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("pad", [False, True])
def test_attention_forward_and_merge(ops, core, rnd, bucket, impl, causal, pad):
    c = _attention_case(core, impl, causal, pad, bucket, clustered=pad)
    B, T, H, R = c["B"], c["T"], c["H"], c["R"]
    sticker, undo = core.sort_buckets(c["buckets"], T)
    q32, v32 = to_bh(c["qk"], H), to_bh(c["v"], H)
    so, slse = core.chunk_attention(q32, v32, sticker, bucket, R, c["spec_o"], c["m_bh"])
    out_x, o_x, lse_ref = core.unsort_and_merge(so, slse, undo, R)                                  # exact fp32 oracle (information)
    want = rnd.forward(q32, v32, sticker, undo, bucket, R, c["spec_o"], c["m_bh"])                  # operand-rounded oracle (asserted)
    mk = None if c["mask"] is None else c["mask"].to(torch.uint8).to(DEV)
    o, lse = ops.lsh_attn_fwd(c["qk"], c["v"], sticker.to(torch.int32).to(DEV).view(B, H, R * T), mk, _gpu_spec(ops, c), H, R, bucket)
    assert report("o_rounds", o.view(B * H, R, T, 64), want["o_rounds"], o_x) <= TOL
    # lse is fp32: absolute 1e-4 against the rounded oracle (the kernel's normaliser is the sum of the bf16 P row, which the exact
    # oracle's differs from by up to 2^-9 relative on a peaked row: printed), except rows whose only target is themselves
    # (lse = self_value ~ -5e4: one fp32 ulp = 4e-3)
    got_lse = lse.cpu().view(B * H, R, T)
    print(f"[parity] lse_rounds: max abs vs rounded oracle {(got_lse - want['lse_rounds']).abs().max():.2e}, vs exact fp32 oracle "
          f"{(got_lse - lse_ref)[lse_ref > -1e4].abs().max():.2e}")
    # (a P element that sits on a bf16 rounding boundary may round the other way on the two sides - exp2 approximations differ in
    # the last fp32 bits - and moves that row's sum by up to 2^-8 of the element's share: a handful of rows per million, bounded
    # by one bf16 ulp of a dominant term; every other row agrees to fp32 accumulation level)
    err = (got_lse - want["lse_rounds"]).abs()
    ok = lse_ref.abs() < 1e4          # (rows that see only themselves: lse = self_value, fp32 spacing 4e-3 .. 8e-3)
    assert (err <= 4e-3 + 2e-7 * lse_ref.abs()).all() and err[ok].pow(2).mean().sqrt().item() <= 1e-4
    assert ((got_lse - lse_ref).abs() <= 4e-3 + 2e-7 * lse_ref.abs()).all()
    out, lse_tot = ops.lsh_merge_fwd(o, lse)
    assert report("merged out", to_bh(out, H), want["out"], out_x) <= TOL
    err = (lse_tot.cpu().view(B * H, T) - want["lse"]).abs()
    ok = want["lse"].abs() < 1e4
    assert (err <= 4e-3 + 2e-7 * want["lse"].abs()).all() and err[ok].pow(2).mean().sqrt().item() <= 1e-4


@pytest.mark.parametrize("bucket", [64, 128])
@pytest.mark.parametrize("impl", ["rp", "hf"])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("pad", [False, True])
def test_attention_backward(ops, core, rnd, bucket, impl, causal, pad):
    c = _attention_case(core, impl, causal, pad, bucket, seed=1)
    B, T, H, R = c["B"], c["T"], c["H"], c["R"]
    dout = torch.randn(B, T, H * 64, device=DEV).bfloat16()
    q32 = to_bh(c["qk"], H).requires_grad_(True)
    v32 = to_bh(c["v"], H).requires_grad_(True)
    res = core.lsh_attention(q32, v32, c["buckets"], bucket, R, c["spec_o"], c["m_bh"])
    (res["out"] * to_bh(dout, H)).sum().backward()                                                  # exact oracle + autograd (information)
    fw = rnd.forward(q32.detach(), v32.detach(), res["sticker"], res["undo"], bucket, R, c["spec_o"], c["m_bh"])
    want_dqk, want_dv = rnd.backward(q32.detach(), v32.detach(), res["sticker"], res["undo"], bucket, R, c["spec_o"], c["m_bh"],
                                     to_bh(dout, H), fw["out"], fw["lse"])                            # operand-rounded oracle (asserted)
    sticker, undo = ops.lsh_sort(c["buckets"].to(torch.int32).to(DEV).view(B, H, R * T), T, R, c["nb"])
    mk = None if c["mask"] is None else c["mask"].to(torch.uint8).to(DEV)
    spec = _gpu_spec(ops, c)
    o, lse_r = ops.lsh_attn_fwd(c["qk"], c["v"], sticker, mk, spec, H, R, bucket)
    out, lse = ops.lsh_merge_fwd(o, lse_r)
    delta = ops.lsh_delta(dout, out, H)
    dqk, dv = ops.lsh_attn_bwd(c["qk"], c["v"], sticker, undo, mk, spec, dout, lse, delta, H, R, bucket)
    assert report("dqk", to_bh(dqk, H), want_dqk, q32.grad) <= TOL
    assert report("dv", to_bh(dv, H), want_dv, v32.grad) <= TOL


@pytest.mark.parametrize("bucket", [64, 128])
@pytest.mark.parametrize("impl,causal,pad", [("rp", True, False), ("rp", False, True), ("hf", True, True)])
def test_attention_large_norm_rows_take_the_exact_two_pass_path(ops, core, rnd, bucket, impl, causal, pad):
    """Rows whose score bound |q| * scale * log2(e) reaches 60 cannot use the single-pass stabiliser (a visible key could
    underflow against the bound): the loader flags such tiles and the softmax group runs the reference's two-pass arithmetic
    (row maximum first).  A third of the tokens get norms of ~400 (bound ~70 and far beyond), the rest stay at ~8, so flagged and
    ordinary tiles alternate inside one launch; forward, merge and backward against the oracle."""
    c = _attention_case(core, impl, causal, pad, bucket, seed=7)
    B, T, H, R = c["B"], c["T"], c["H"], c["R"]
    big = (torch.arange(T, device=DEV) % 3 == 0).view(1, T, 1)
    c["qk"] = torch.where(big, c["qk"].float() * 50.0, c["qk"].float()).bfloat16()
    dout = torch.randn(B, T, H * 64, device=DEV).bfloat16()
    q32 = to_bh(c["qk"], H).requires_grad_(True)
    v32 = to_bh(c["v"], H).requires_grad_(True)
    res = core.lsh_attention(q32, v32, c["buckets"], bucket, R, c["spec_o"], c["m_bh"])
    (res["out"] * to_bh(dout, H)).sum().backward()
    sticker, undo = ops.lsh_sort(c["buckets"].to(torch.int32).to(DEV).view(B, H, R * T), T, R, c["nb"])
    mk = None if c["mask"] is None else c["mask"].to(torch.uint8).to(DEV)
    spec = _gpu_spec(ops, c)
    o, lse_r = ops.lsh_attn_fwd(c["qk"], c["v"], sticker, mk, spec, H, R, bucket)
    out, lse = ops.lsh_merge_fwd(o, lse_r)
    assert torch.isfinite(out.float()).all() and torch.isfinite(lse).all()
    fw = rnd.forward(q32.detach(), v32.detach(), res["sticker"], res["undo"], bucket, R, c["spec_o"], c["m_bh"])
    want_dqk, want_dv = rnd.backward(q32.detach(), v32.detach(), res["sticker"], res["undo"], bucket, R, c["spec_o"], c["m_bh"],
                                     to_bh(dout, H), fw["out"], fw["lse"])
    # the kernel switches WHOLE tiles to the exact arithmetic, the rounded oracle single rows: rows of a flagged tile whose own
    # bound is small round P against a different stabiliser on the two sides - independent bf16 roundings of P, 2e-3 at most
    assert report("out (exact path)", to_bh(out, H), fw["out"], res["out"]) <= 2 * TOL
    delta = ops.lsh_delta(dout, out, H)
    dqk, dv = ops.lsh_attn_bwd(c["qk"], c["v"], sticker, undo, mk, spec, dout, lse, delta, H, R, bucket)
    assert report("dv (exact path)", to_bh(dv, H), want_dv, v32.grad) <= 2 * TOL
    # d/dqk of near-one-hot rows is a difference of large terms (delta - dP cancels to rounding level): the kernel's bf16 `out`
    # differs from the oracle's by rare roundings, which this cancellation amplifies
    assert report("dqk (exact path)", to_bh(dqk, H), want_dqk, q32.grad) <= 5 * TOL


def test_attention_first_chunk_looks_back_at_last_chunk_of_previous_round(ops, core, rnd):
    """Look-one-back wraps: chunk 0 of round r sees the last chunk of round r-1, and chunk 0 of round 0 the very last
    chunk (rp R6).  One distinctive value row in the last chunk must reach queries of the first chunk."""
    B, T, H, R, bucket = 1, 256, 1, 2, 64
    qk = torch.randn(B, T, 64, device=DEV).bfloat16()
    v = torch.zeros(B, T, 64, device=DEV).bfloat16()
    nb = T // bucket
    buckets = (torch.arange(T) // bucket).view(1, 1, T) + nb * torch.arange(R).view(1, R, 1)     # identity sort
    buckets = buckets.view(1, R * T)
    v[0, T - 1] = 7.0      # last token of the last chunk
    sticker, undo = core.sort_buckets(buckets, T)
    spec_o = core.LSHSpec.reformer_pytorch(64, False)
    ref = core.lsh_attention(to_bh(qk, 1), to_bh(v, 1), buckets, bucket, R, spec_o)
    o, lse = ops.lsh_attn_fwd(qk, v, sticker.to(torch.int32).to(DEV).view(1, 1, -1), None, ops.LSHSpec.reformer_pytorch(64, False), 1, R, bucket)
    out, _ = ops.lsh_merge_fwd(o, lse)
    assert ref["out"][0, :bucket].abs().min() > 0        # oracle: first chunk sees token T-1
    want = rnd.forward(to_bh(qk, 1), to_bh(v, 1), sticker, undo, bucket, R, spec_o)
    assert report("look-back out", to_bh(out, 1), want["out"], ref["out"]) <= TOL


def test_attention_sizes_of_baseline_configs(ops, core):
    """Full-size property check (no oracle): merge weights sum to one, outputs finite, and a second run is bit-identical."""
    B, T, H, R, bucket = 4, 1024, 8, 8, 64
    torch.manual_seed(3)
    qk = torch.randn(B, T, H * 64, device=DEV).bfloat16()
    v = torch.ones(B, T, H * 64, device=DEV).bfloat16()      # constant V: every convex combination returns exactly 1
    rot = torch.randn(1, 64, R, (T // bucket) // 2, device=DEV)
    buckets = ops.lsh_hash(qk, rot, H, R, T // bucket)
    sticker, undo = ops.lsh_sort(buckets, T, R, T // bucket)
    idx = torch.arange(R * T, device=DEV, dtype=torch.int32).expand(B, H, -1)
    assert torch.equal(torch.gather(undo, 2, sticker.long()), idx)                                  # undo inverts sticker
    sorted_b = torch.gather(buckets, 2, sticker.long())
    assert bool((sorted_b[..., 1:] >= sorted_b[..., :-1]).all())                                    # sortedness
    spec = ops.LSHSpec.reformer_pytorch(64, True)
    o1, l1 = ops.lsh_attn_fwd(qk, v, sticker, None, spec, H, R, bucket)
    o2, l2 = ops.lsh_attn_fwd(qk, v, sticker, None, spec, H, R, bucket)
    assert torch.equal(o1, o2) and torch.equal(l1, l2)                                              # deterministic
    out, lse = ops.lsh_merge_fwd(o1, l1)
    assert torch.isfinite(out.float()).all() and (out.float() - 1).abs().max().item() <= 2 ** -7    # rows of P sum to 1


@pytest.mark.parametrize("B,T,H,R,causal,pad", [(3, 1024, 8, 8, True, 0), (5, 256, 8, 8, False, 40), (2, 128, 2, 1, True, 0), (1, 2048, 4, 3, True, 300),
                                               (37, 128, 4, 2, False, 17)])
def test_attention_bucket64_kernels_agree(ops, core, B, T, H, R, causal, pad):
    """The three bucket-64 forward kernels (paired-chunk = the default, 128-query tiles, block-streaming) are independent
    implementations of the same arithmetic: per-round outputs and log-sum-exps must agree to bf16 / fp32 rounding.  Shapes cover
    runs of a CTA that start inside a (batch, head) row, rows of a single tile (T = 128, R = 1), more CTAs than tiles and fewer."""
    import ctypes
    from reformer_tts_b200 import _lib
    lib = _lib.load()
    lib.rtts_debug_set_fwd_kernel.argtypes = [ctypes.c_int]
    torch.manual_seed(11)
    qkv = torch.randn(B, T, 2 * H * 64, device=DEV).bfloat16()
    qk, v = qkv[..., :H * 64], qkv[..., H * 64:]
    nb = T // 64
    rot = torch.randn(1, 64, R, nb // 2, device=DEV)
    mask = None
    if pad:
        mask = torch.ones(B, T, dtype=torch.uint8, device=DEV)
        mask[:, T - pad:] = 0
    spec = ops.LSHSpec.reformer_pytorch(64, causal)
    buckets, sumsq = ops.lsh_hash(qk, rot, H, R, nb, return_sumsq=True)
    sticker, _ = ops.lsh_sort(buckets, T, R, nb)
    res = []
    try:
        for which in (0, 1, 2):
            lib.rtts_debug_set_fwd_kernel(which)
            o, lse = ops.lsh_attn_fwd(qk, v, sticker, mask, spec, H, R, 64, sumsq=sumsq)
            torch.cuda.synchronize()
            res.append((o.float(), lse.clone()))
    finally:
        lib.rtts_debug_set_fwd_kernel(0)
    for which, name in ((1, "tile"), (2, "block")):
        # (the kernels form the stabiliser with differently rounded arithmetic, so a P value near a bf16 rounding boundary can fall
        # on either side: agreement is to bf16 OUTPUT rounding, 2^-9 / sqrt(3) = 1.1e-3 if every element differed by half an ulp;
        # each kernel alone is held to TOL against the rounded oracle by the tests above)
        assert rel_l2(res[0][0], res[which][0]) <= 2e-3, name
        assert (res[0][1] - res[which][1]).abs().max().item() <= 4e-3, name      # fp32 spacing at 5e4 is 4e-3


@pytest.mark.parametrize("causal", [False, True])
def test_attention_full_size_with_heavy_padding(ops, core, rnd, causal):
    """BASELINE.json config 2 decoder shape (B=20, T=1024, 8 heads, 8 rounds, bucket 64) with 224 padded positions per sequence
    (800 mel frames padded to 1024).  22 % of the rows see only themselves, which makes the epilogue the slowest role: this is the
    configuration in which a consumer waiting on a barrier that may run two phases ahead deadlocks.  Properties: finishes, is
    deterministic, a padded query returns exactly its own value row (rp R8: everything but the self column masked), and one
    (batch, head) slice equals the oracle."""
    B, T, H, R, bucket, pad = 20, 1024, 8, 8, 64, 224
    torch.manual_seed(5)
    qk = torch.randn(B, T, H * 64, device=DEV).bfloat16()
    v = torch.randn(B, T, H * 64, device=DEV).bfloat16()
    rot = torch.randn(1, 64, R, (T // bucket) // 2, device=DEV)
    mask = torch.ones(B, T, dtype=torch.uint8, device=DEV)
    mask[:, T - pad:] = 0
    buckets = ops.lsh_hash(qk, rot, H, R, T // bucket)
    sticker, undo = ops.lsh_sort(buckets, T, R, T // bucket)
    spec = ops.LSHSpec.reformer_pytorch(64, causal)
    o1, l1 = ops.lsh_attn_fwd(qk, v, sticker, mask, spec, H, R, bucket)
    o2, l2 = ops.lsh_attn_fwd(qk, v, sticker, mask, spec, H, R, bucket)
    torch.cuda.synchronize()
    assert torch.equal(o1, o2) and torch.equal(l1, l2)
    assert torch.isfinite(o1.float()).all() and torch.isfinite(l1).all()
    v_bh = v.view(B, T, H, 64).permute(0, 2, 1, 3)                          # [B,H,T,64]
    assert torch.equal(o1[:, :, :, T - pad:], v_bh[:, :, None, T - pad:].expand(B, H, R, pad, 64))
    # one (batch, head) slice against the oracle, fed our bucket ids
    b, h = 7, 3
    bk = buckets[b, h].cpu().long().view(1, R * T)
    st, ud = core.sort_buckets(bk, T)
    q1, v1 = qk[b:b + 1, :, h * 64:(h + 1) * 64], v[b:b + 1, :, h * 64:(h + 1) * 64]
    so, slse = core.chunk_attention(to_bh(q1, 1), to_bh(v1, 1), st, bucket, R, core.LSHSpec.reformer_pytorch(64, causal),
                                    mask[b:b + 1].bool().cpu())
    out_ref, o_ref, lse_ref = core.unsort_and_merge(so, slse, ud, R)
    want = rnd.forward(to_bh(q1, 1), to_bh(v1, 1), st, ud, bucket, R, core.LSHSpec.reformer_pytorch(64, causal), mask[b:b + 1].bool().cpu())
    assert report("full-size o_rounds slice", o1[b, h], want["o_rounds"][0], o_ref[0]) <= TOL
    assert ((l1[b, h].cpu() - want["lse_rounds"][0]).abs() <= 4e-3 + 2e-7 * lse_ref[0].abs()).all()


@pytest.mark.parametrize("name,B,T,H,R,bucket,causal,pad,impl", [
    ("cfg2-decoder", 20, 1024, 8, 8, 64, True, 224, "rp"),        # config/bucket-size-64-18-06.yml decoder layer, 800 frames padded to 1024
    ("cfg2-encoder", 20, 256, 8, 8, 64, False, 56, "rp"),         # its encoder layer, 200 phonemes padded to 256
    ("cfg1-decoder", 4, 1024, 8, 8, 128, True, 224, "rp"),        # config/baseline.yml / depth-3-15-06.yml decoder layer (bucket 128)
    ("cfg3-decoder-hf", 12, 1024, 8, 8, 128, True, 224, "hf"),    # config/huggingface-lsh.yml decoder layer (HF semantics, 16 buckets)
    ("cfg5-sweep-4k", 1, 4096, 8, 4, 64, True, 0, "rp"),          # long-sequence sweep, 4096 positions, 4 rounds
])
def test_attention_config_shapes_forward_and_backward_slices(ops, core, rnd, name, B, T, H, R, bucket, causal, pad, impl):
    """Full-size launches at the shapes of BASELINE.json's configs (hash -> sort -> attention -> merge -> backward on the GPU), then
    two (batch, head) slices of every result against the operand-rounded oracle fed OUR bucket ids: per-round outputs, merged
    output, and both gradients at 1e-3."""
    torch.manual_seed(len(name) + T)
    qk = torch.randn(B, T, H * 64, device=DEV).bfloat16()
    v = torch.randn(B, T, H * 64, device=DEV).bfloat16()
    dout = torch.randn(B, T, H * 64, device=DEV).bfloat16()
    nb = T // bucket if impl == "rp" else 2 ** ((2 * (T // bucket)).bit_length() - 1)
    rot = torch.randn(1 if impl == "rp" else H, 64, R, nb // 2, device=DEV)
    mask = None
    if pad:
        mask = torch.ones(B, T, dtype=torch.uint8, device=DEV)
        mask[:, T - pad:] = 0
        mask[0, T // 2:] = 0
    use_pad_bucket = impl == "hf" and pad > 0
    buckets = ops.lsh_hash(qk, rot, H, R, nb, mask if use_pad_bucket else None, use_pad_bucket)
    ids = nb + 1 if use_pad_bucket else nb
    sticker, undo = ops.lsh_sort(buckets, T, R, ids)
    spec = (ops.LSHSpec.reformer_pytorch if impl == "rp" else ops.LSHSpec.huggingface)(64, causal)
    spec_o = (core.LSHSpec.reformer_pytorch if impl == "rp" else core.LSHSpec.huggingface)(64, causal)
    o, lse_r = ops.lsh_attn_fwd(qk, v, sticker, mask, spec, H, R, bucket)
    out, lse = ops.lsh_merge_fwd(o, lse_r)
    delta = ops.lsh_delta(dout, out, H)
    dqk, dv = ops.lsh_attn_bwd(qk, v, sticker, undo, mask, spec, dout, lse, delta, H, R, bucket)
    torch.cuda.synchronize()
    assert torch.isfinite(dqk.float()).all() and torch.isfinite(dv.float()).all()
    for b, h in [(0, 0), (B - 1, H - 3)]:
        sl = lambda a: a[b:b + 1, :, h * 64:(h + 1) * 64]
        q1, v1, d1 = to_bh(sl(qk), 1), to_bh(sl(v), 1), to_bh(sl(dout), 1)
        bk = buckets[b, h].cpu().long().view(1, R * T)
        st, ud = core.sort_buckets(bk, T)
        assert torch.equal(st, sticker[b, h].cpu().long().view(1, -1))
        m1 = None if mask is None else mask[b:b + 1].bool().cpu()
        fw = rnd.forward(q1, v1, st, ud, bucket, R, spec_o, m1)
        assert report(f"{name} o_rounds[{b},{h}]", o[b, h], fw["o_rounds"][0]) <= TOL
        assert report(f"{name} out[{b},{h}]", to_bh(sl(out), 1), fw["out"]) <= TOL
        assert (lse[b, h].cpu() - fw["lse"][0]).abs().max().item() <= 4e-3 + 2e-7 * fw["lse"].abs().max().item()
        gq, gv = rnd.backward(q1, v1, st, ud, bucket, R, spec_o, m1, d1, fw["out"], fw["lse"])
        assert report(f"{name} dqk[{b},{h}]", to_bh(sl(dqk), 1), gq) <= TOL
        assert report(f"{name} dv[{b},{h}]", to_bh(sl(dv), 1), gv) <= TOL


# ---------------------------------------------------------------------------------------------- GEMM / row-wise kernels
@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (256, 384, 512), (2048, 2048, 512), (1024, 512, 2048)])
def test_gemm_layouts_and_epilogues(ops, m, n, k):
    torch.manual_seed(m + n + k)
    a = torch.randn(m, k, device=DEV).bfloat16()
    b = torch.randn(n, k, device=DEV).bfloat16()
    ref = a.double() @ b.double().t()
    at, bt = a.t().contiguous(), b.t().contiguous()
    for amn, bmn in [(False, False), (True, False), (False, True), (True, True)]:
        c = ops.gemm(at if amn else a, bt if bmn else b, a_mn_major=amn, b_mn_major=bmn)
        assert rel_l2(c, ref) <= 1e-5, (amn, bmn)
    bias = torch.randn(n, device=DEV)
    gate = torch.randn(m, n, device=DEV).bfloat16()
    assert rel_l2(ops.gemm(a, b, bias=bias, relu=True, out_dtype=torch.bfloat16), bf16r(torch.relu(ref + bias.double()).float())) <= TOL
    cs = torch.zeros(n, device=DEV)
    gated = ops.gemm(a, b, gate=gate, colsum=cs, out_dtype=torch.bfloat16)
    want = ref * (gate.double() > 0)
    assert rel_l2(gated, bf16r(want.float())) <= TOL and rel_l2(cs, want.sum(0)) <= 1e-4
    if k >= 512:
        acc = torch.full((m, n), 2.0, device=DEV)
        ops.gemm(at, bt, a_mn_major=True, b_mn_major=True, out=acc, accumulate=True, split_k=4)
        assert rel_l2(acc, ref + 2) <= 1e-5


@pytest.mark.parametrize("sub", [False, True])
def test_gemm_residual_epilogue(ops, sub):
    """C = resid +/- (A . B^T + bias) written from the epilogue (the reversible residual of the FFN / cross-attention blocks:
    y = x + f(.) in the forward, x = y - f(.) in the recompute), also with C aliasing resid."""
    torch.manual_seed(11)
    m, n, k = 512, 256, 384
    a = torch.randn(m, k, device=DEV).bfloat16()
    b = torch.randn(n, k, device=DEV).bfloat16()
    bias = torch.randn(n, device=DEV)
    resid = torch.randn(m, n, device=DEV) * 10
    want = a.float() @ b.float().t() + bias
    want = resid - want if sub else resid + want
    got = ops.gemm(a, b, bias=bias, resid=resid, resid_sub=sub)
    assert rel_l2(got, want) <= TOL_FP32
    inplace = resid.clone()
    got2 = ops.gemm(a, b, bias=bias, resid=inplace, resid_sub=sub, out=inplace)
    assert got2.data_ptr() == inplace.data_ptr() and torch.equal(got2, got)
    # with the inverted-dropout keep mask in the same epilogue (rtts_gemm_bf16_dropout), and on its own
    keep = (torch.rand(m, n, device=DEV) < 0.85).to(torch.uint8)
    scale = 1.0 / 0.85
    core = (a.float() @ b.float().t() + bias) * keep * scale
    got3 = ops.gemm(a, b, bias=bias, resid=resid, resid_sub=sub, keep_mask=keep, keep_scale=scale)
    assert rel_l2(got3, resid - core if sub else resid + core) <= TOL_FP32
    got4 = ops.gemm(a, b, bias=bias, keep_mask=keep, keep_scale=scale)
    assert rel_l2(got4, core) <= TOL_FP32 and bool((got4[keep == 0] == 0).all())


def test_gemm_rejects_bad_shapes(ops):
    a = torch.randn(100, 64, device=DEV).bfloat16()
    b = torch.randn(128, 64, device=DEV).bfloat16()
    with pytest.raises(RuntimeError, match="multiples of 128"):
        ops.gemm(a, b)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.gemm(a.cpu(), b.cpu())


@pytest.mark.parametrize("dim", [128, 512])
def test_layernorm_forward_backward(ops, dim):
    torch.manual_seed(dim)
    x = torch.randn(4096 + 8, dim, device=DEV) * 2 + 0.5
    g, bt = torch.randn(dim, device=DEV), torch.randn(dim, device=DEV)
    y, mean, rstd = ops.layernorm_fwd(x, g, bt)
    xr, gr, br = x.clone().requires_grad_(True), g.clone().requires_grad_(True), bt.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (dim,), gr, br, 1e-5)
    assert rel_l2(y, bf16r(yr.detach())) <= TOL
    dy = torch.randn_like(x)
    yr.backward(dy)
    dg, db = torch.ones(dim, device=DEV), torch.ones(dim, device=DEV)       # accumulate semantics: += on top of 1
    dx = ops.layernorm_bwd(dy, x, g, mean, rstd, dg, db)
    assert rel_l2(dx, xr.grad) <= 1e-5 and rel_l2(dg - 1, gr.grad) <= 1e-4 and rel_l2(db - 1, br.grad) <= 1e-4
    # accumulate-onto variant (the reversible blocks' dx2 += df): a pending GradAccumRequest is added in the same kernel
    from reformer_tts_b200.residual import GradAccumRequest
    base = torch.randn_like(x)
    dg2, db2 = torch.zeros(dim, device=DEV), torch.zeros(dim, device=DEV)
    with GradAccumRequest(base) as req:
        dx_acc = ops.layernorm_bwd(dy, x, g, mean, rstd, dg2, db2, accumulate_request=True)
    assert req.consumed and rel_l2(dx_acc, base + xr.grad) <= 1e-5
    with GradAccumRequest(base[:, : dim // 2]) as req:      # shape / contiguity mismatch: not taken, plain result
        dx_plain = ops.layernorm_bwd(dy, x, g, mean, rstd, dg2, db2, accumulate_request=True)
    assert not req.consumed and torch.equal(dx_plain, dx)


def test_cast_colsum_and_delta(ops):
    x = torch.randn(1000, 512, device=DEV)
    cs = torch.zeros(512, device=DEV)
    y = ops.cast_bf16_colsum(x, cs)
    assert torch.equal(y, x.bfloat16()) and rel_l2(cs, x.sum(0)) <= 1e-5
    # inverted dropout applied first (backward of rtts_gemm_bf16_dropout)
    keep = (torch.rand(1000, 512, device=DEV) < 0.8).to(torch.uint8)
    cs2 = torch.zeros(512, device=DEV)
    y2 = ops.cast_bf16_colsum(x, cs2, keep_mask=keep, keep_scale=1.25)
    want2 = x * keep * 1.25
    assert torch.equal(y2, want2.bfloat16()) and rel_l2(cs2, want2.sum(0)) <= 1e-5
    a = torch.randn(3, 256, 512, device=DEV).bfloat16()
    o = torch.randn(3, 256, 512, device=DEV).bfloat16()
    want = (a.float() * o.float()).view(3, 256, 8, 64).sum(-1).permute(0, 2, 1)
    assert rel_l2(ops.lsh_delta(a, o, 8), want) <= 1e-5


@pytest.mark.parametrize("rows,cols,ld", [(20480, 512, 512), (5120, 1024, 1024), (777, 64, 96), (3, 8, 8)])
def test_colsum_bf16(ops, rows, cols, ld):
    """Bias gradients of bf16 gradient matrices: fp32 column sums accumulated into the target (strided rows allowed)."""
    torch.manual_seed(2)
    buf = torch.randn(rows, ld, device=DEV).bfloat16()
    x = buf[:, :cols]
    acc = torch.full((cols,), 0.5, device=DEV)
    ops.colsum_bf16(x, acc)
    want = x.double().sum(0) + 0.5
    assert (acc.double() - want).abs().max().item() <= 1e-5 * max(1.0, x.double().abs().sum(0).max().item())


def _xattn_drop_mask(seed: int, B, H, T, S, p_drop):
    """The kernels' counter-based dropout mask (csrc/xattn.cu drop_hash), restated with 64-bit integer tensors: keep [B,H,T,S] bool."""
    M = 0xFFFFFFFF
    row = torch.arange(B * H * T, dtype=torch.int64).view(B, H, T, 1)
    pair = torch.arange(S // 2, dtype=torch.int64).view(1, 1, 1, S // 2)
    x = ((row * (S // 2) + pair) & M) * 0x9E3779B1 + (seed & M) & M
    x = x & M
    x ^= x >> 15
    x = (x * 0x85EBCA77) & M
    x ^= x >> 13
    x = (x * 0xC2B2AE3D + ((seed >> 32) & M)) & M
    x ^= x >> 16
    thr = int(p_drop * 65536.0)
    keep = torch.stack(((x & 0xFFFF) >= thr, (x >> 16) >= thr), dim=-1).view(B, H, T, S)
    return keep


@pytest.mark.parametrize("B,T,S,H,pad,p_drop", [(2, 256, 256, 2, True, 0.0), (3, 128, 128, 8, False, 0.15), (2, 1024, 256, 8, True, 0.15), (1, 128, 64, 1, True, 0.0)])
def test_xattn_forward_backward(ops, B, T, S, H, pad, p_drop):
    """Dense decoder -> encoder attention core (nn.MultiheadAttention's softmax(QK^T / sqrt(dh)) V with key padding mask and
    probability dropout, ref:reformer_tts/model/reformer.py:161-186) against the same arithmetic in fp32 with the kernels' operand
    roundings (bf16 q / k / v / dout, P and dS rounded to bf16 where the kernels store them) and the regenerated dropout mask."""
    torch.manual_seed(9)
    D = H * 64
    q = torch.randn(B, T, D, device=DEV).bfloat16()
    kv = torch.randn(B, S, 2 * D, device=DEV).bfloat16()
    dout = torch.randn(B, T, D, device=DEV).bfloat16()
    keep = None
    if pad:
        keep = torch.ones(B, S, dtype=torch.uint8, device=DEV)
        for b in range(B):
            keep[b, S - 7 - 13 * b:] = 0
    seed = torch.tensor([0x1234567_89ABCDE], dtype=torch.int64, device=DEV)
    scale = 0.125
    k, v = kv[..., :D], kv[..., D:]
    out, lse = ops.xattn_fwd(q, k, v, keep, H, scale, p_drop, seed)
    delta = ops.lsh_delta(dout, out, H)
    dq, dkv = ops.xattn_bwd(q, k, v, keep, H, scale, p_drop, seed, dout, lse, delta)
    torch.cuda.synchronize()

    # exact reference (CPU, fp64, autograd) ...
    qf = q.double().cpu().view(B, T, H, 64).transpose(1, 2).requires_grad_(True)
    kf = k.double().cpu().reshape(B, S, H, 64).transpose(1, 2).requires_grad_(True)
    vf = v.double().cpu().reshape(B, S, H, 64).transpose(1, 2).requires_grad_(True)
    s = qf @ kf.transpose(-1, -2) * scale
    if keep is not None:
        s = s.masked_fill(~keep.bool().cpu()[:, None, None, :], float("-inf"))
    prob = torch.softmax(s, dim=-1)
    dmask = torch.ones_like(prob)
    if p_drop > 0:
        dmask = _xattn_drop_mask(int(seed.item()), B, H, T, S, p_drop).double() / (1.0 - p_drop)
    ref = (prob * dmask) @ vf
    go = dout.double().cpu().view(B, T, H, 64).transpose(1, 2)
    ref.backward(go)
    # ... and the same arithmetic with a bf16 rounding exactly where the kernels store a bf16 operand: the un-normalised
    # exponentials (forward), D o P and dS (backward), the outputs
    with torch.no_grad():
        r16 = lambda x: x.float().bfloat16().double()
        sd = s.detach()
        e = torch.exp(sd - sd.amax(-1, keepdim=True))
        keep01 = (dmask > 0).double()
        out_r = r16((r16(e) * keep01) @ vf / e.sum(-1, keepdim=True) / (1.0 - p_drop))
        dp = go @ vf.transpose(-1, -2)
        dl = (go * bf16r(out.float()).double().cpu().view(B, T, H, 64).transpose(1, 2)).sum(-1, keepdim=True)       # delta from the stored output
        pr = prob.detach()
        ds_r = r16(pr * (dmask * dp - dl) * scale)
        dq_r, dk_r, dv_r = r16(ds_r @ kf), ds_r.transpose(-1, -2) @ qf, r16(pr * dmask).transpose(-1, -2) @ go
    lse_ref = torch.logsumexp(s.detach(), dim=-1) / 0.6931471805599453      # log2 domain
    bh = lambda x: x.double().cpu().view(B, -1, H, 64).transpose(1, 2)
    assert report("xattn out", bh(out), out_r, ref.detach()) <= TOL
    assert (lse.double().cpu() - lse_ref).abs().max().item() <= 1e-3
    assert report("xattn dq", bh(dq), dq_r, qf.grad) <= TOL
    assert report("xattn dk", bh(dkv[..., :D]), dk_r, kf.grad) <= TOL
    assert report("xattn dv", bh(dkv[..., D:]), dv_r, vf.grad) <= TOL
