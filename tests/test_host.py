"""CPU-side checks: the C-ABI library loads and exports every symbol include/rtts_b200.h declares, the host mirror keeps
the reference's constructor / state-dict contract, the reversible algebra (pure PyTorch) matches the oracle and the
reference's golden gradients, and the data-parallel gradient averaging works over gloo with world_size 2."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_library_loads_and_exports_every_declared_symbol():
    from reformer_tts_b200 import _lib
    header = open(os.path.join(ROOT, "include", "rtts_b200.h")).read()
    declared = set(re.findall(r"\b(rtts_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 12
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared - {"rtts_last_error", "rtts_abi_version", "rtts_build_id"} == set(_lib.SIGNATURES), "ctypes table out of sync with the header"
    assert lib.rtts_abi_version() == 1
    # the binary was built from exactly the sources that are checked out (a stale prebuilt .so must not pass for HEAD)
    from reformer_tts_b200.csrc.build import source_hash
    assert _lib.build_id() == source_hash(), "libreformer_b200.so is stale: run python reformer_tts_b200/csrc/build.py"
    # argument validation happens before any CUDA call: a bad head size is reported, not launched
    rc = lib.rtts_lsh_hash(None, 0, None, 1, None, 0, None, None, 1, 128, 1, 64, 1, 2, None)
    assert rc != 0 and b"null pointer" in lib.rtts_last_error()


def test_hash_tc_shape_rules_through_the_c_abi():
    """Host-only entry points of the tensor-pipe hash (no CUDA call): which shapes it takes and how much workspace it wants."""
    from reformer_tts_b200 import _lib
    lib = _lib.load()
    sup = lib.rtts_lsh_hash_tc_supported
    # (T, dh, R, n_buckets): the four training configs and the sweep lengths
    assert sup(1024, 64, 8, 16) and sup(256, 64, 8, 4) and sup(1024, 64, 8, 8) and sup(256, 64, 8, 8)
    assert sup(16384, 64, 4, 256)          # 512 projections: two launches of two rounds (256 columns each)
    assert not sup(384, 64, 8, 6)          # 24 projections: not a multiple of 16 -> fp32 FMA kernel
    assert not sup(1000, 64, 8, 16) and not sup(1024, 32, 8, 16) and not sup(1024, 64, 8, 1024)
    ws = lib.rtts_lsh_hash_tc_workspace_bytes
    assert ws(1, 8, 16) == 3 * 64 * 64 * 2 and ws(8, 8, 16) == 8 * 3 * 64 * 64 * 2
    assert ws(1, 4, 256) == 3 * 256 * 64 * 2      # one group of two rounds at a time
    rc = lib.rtts_lsh_hash_tc(None, 0, None, 1, None, 0, None, None, None, 1, 128, 1, 64, 8, 4, None)
    assert rc != 0 and b"null pointer" in lib.rtts_last_error()


def test_xattn_and_colsum_argument_rules_through_the_c_abi():
    """Argument validation of the dense cross-attention core and of the bf16 column sum happens before any CUDA call: null pointers,
    sequence lengths the kernels do not take (T not a multiple of 128; S not a multiple of 32 / 64 or beyond 256), head sizes other
    than 64 and dropout without a seed word are reported, not launched.  ``ops.xattn_supported`` states the same rules for the host."""
    import ctypes
    from reformer_tts_b200 import _lib, ops
    lib = _lib.load()
    buf = ctypes.create_string_buffer(4096)
    p = ctypes.cast(buf, ctypes.c_void_p)      # 16-byte aligned host memory: never dereferenced, the calls below fail in validation
    p = ctypes.c_void_p((p.value + 15) & ~15)

    def fwd(T=1024, S=256, dh=64, p_drop=0.0, seed=None, q=p):
        return lib.rtts_xattn_fwd(q, 512, p, p, 1024, None, 0.125, p_drop, seed, p, 512, p, 2, T, S, 8, dh, None)

    def bwd(T=1024, S=256, dh=64):
        return lib.rtts_xattn_bwd(p, 512, p, p, 1024, None, 0.125, 0.0, None, p, 512, p, p, p, 512, p, p, 1024, 2, T, S, 8, dh, None)

    assert fwd(q=None) != 0 and b"null pointer" in lib.rtts_last_error()
    assert fwd(T=1000) != 0 and b"multiple of 128" in lib.rtts_last_error()
    assert fwd(S=320) != 0 and fwd(S=48) != 0 and b"S=" in lib.rtts_last_error()
    assert fwd(dh=32) != 0 and b"head size" in lib.rtts_last_error()
    assert fwd(p_drop=0.15) != 0 and b"seed" in lib.rtts_last_error()
    assert bwd(S=96) != 0 and b"multiple of 64" in lib.rtts_last_error()
    assert bwd(T=64) != 0 and bwd(dh=128) != 0
    assert ops.xattn_supported(1024, 256, 64) and ops.xattn_supported(128, 64, 64)
    assert not ops.xattn_supported(1000, 256, 64) and not ops.xattn_supported(1024, 288, 64) and not ops.xattn_supported(1024, 256, 32)
    assert lib.rtts_colsum_bf16(None, 8, p, 4, 8, None) != 0 and b"null pointer" in lib.rtts_last_error()
    assert lib.rtts_colsum_bf16(p, 8, p, 4, 12, None) != 0 and b"multiples of 8" in lib.rtts_last_error()


def test_product_modules_refuse_cpu_tensors():
    from reformer_tts_b200.lsh_attention import LSHSelfAttention
    with pytest.raises(RuntimeError, match="no CPU path"):
        LSHSelfAttention(128, heads=2)(torch.randn(1, 128, 128))


def test_unsupported_kwargs_raise_at_construction():
    from reformer_tts_b200.lsh_attention import LSHSelfAttention
    from reformer_tts_b200.model import FeedForward, LSHSelfAttentionWrapper
    for bad in (dict(use_full_attn=True), dict(num_mem_kv=4), dict(dropout=0.1), dict(random_rotations_per_head=True),
                dict(attend_across_buckets=False), dict(allow_duplicate_attention=False), dict(one_value_head=True),
                dict(return_attn=True), dict(attn_chunks=2), dict(bucket_size=32), dict(heads=4)):
        with pytest.raises(NotImplementedError):
            LSHSelfAttention(512, **bad)
    with pytest.raises(NotImplementedError):
        FeedForward(512, 2048, dropout=0.1)
    with pytest.raises(ValueError):
        LSHSelfAttentionWrapper(512, causal=False, implementation="other", heads=8)


def test_state_dict_keys_match_reference_model():
    """Checkpoint compatibility: parameter names / shapes equal those of the reference ReformerTTS (HF variant, instantiated
    from /root/reference by tests/golden/make_golden.py)."""
    from reformer_tts_b200.model import ReformerTTS, config as C
    want = json.load(open(os.path.join(GOLDEN, "state_dict_keys_hf.json")))
    ours = {k: list(v.shape) for k, v in ReformerTTS(**C.reference_model_kwargs("huggingface-lsh")).state_dict().items()}
    # transformers 5.x registers its mask constants as non-persistent buffers, so they are not in the golden list either
    assert ours == want
    rp = ReformerTTS(**C.reference_model_kwargs("baseline")).state_dict()
    base = "enc.reformer.layers.blocks.0.f.net.fn.layer."
    assert {k[len(base):] for k in rp if k.startswith(base)} == {"toqk.weight", "tov.weight", "to_out.weight", "to_out.bias"}
    assert "dec.reformer.layers.blocks.4.f.net.fn.fn.net.3.bias" in rp and "dec.reformer.layers.blocks.2.f.net.fn.layer.in_proj_weight" in rp


def test_config_loader_applies_reference_defaults():
    from reformer_tts_b200.model import config as C
    kw = C.reference_model_kwargs("bucket-size-64-18-06")
    assert kw["enc_reformer_kwargs"]["depth"] == 6 and kw["dec_reformer_kwargs"]["self_attn_kwargs"]["bucket_size"] == 64
    assert kw["enc_reformer_kwargs"]["attn_kwargs"]["post_attn_dropout"] == 0.15 and kw["pad_base"] == 256
    assert C.reference_model_kwargs("baseline")["dec_reformer_kwargs"]["self_attn_kwargs"]["bucket_size"] == 128
    with pytest.raises(KeyError):
        C.model_kwargs({"num_mel_coeffs": 80, "dict_size": 76, "no_such_key": 1})


def _mlp(d, seed):
    torch.manual_seed(seed)
    return nn.Sequential(nn.LayerNorm(d), nn.Linear(d, 2 * d), nn.ReLU(), nn.Linear(2 * d, d))


def test_reversible_mirror_matches_reference_golden_gradients():
    """The product's reversible.py is pure PyTorch, so it is checked on CPU against the gradients the reference's own file
    produced (tests/golden/reversible_ref.npz)."""
    from reformer_tts_b200.model.reversible import ReversibleBlock, ReversibleHalfResidual, ReversibleSequence, ReversibleSwap
    z = np.load(os.path.join(GOLDEN, "reversible_ref.npz"))
    nets = [_mlp(16, 10 + i) for i in range(7)]
    blocks = nn.ModuleList([ReversibleBlock(nets[0], nets[1]), ReversibleBlock(nets[2], nets[3]), ReversibleHalfResidual(nets[4]),
                            ReversibleSwap(), ReversibleHalfResidual(nets[5]), ReversibleSwap()])
    seq = ReversibleSequence(blocks).train()
    x = torch.from_numpy(z["x"]).requires_grad_(True)
    y = seq(x, kwargs_list=[{}] * 6)
    (y * torch.from_numpy(z["w"])).sum().backward()
    assert (y.detach() - torch.from_numpy(z["y"])).abs().max().item() <= 1e-6
    assert (x.grad - torch.from_numpy(z["dx"])).abs().max().item() <= 1e-5
    for i, n in enumerate(nets[:6]):
        for k, p in n.named_parameters():
            assert (p.grad - torch.from_numpy(z[f"g{i}_{k}"])).abs().max().item() <= 1e-5, (i, k)
    # public concatenated-tensor interface (forward / backward_pass) of a single block
    blk = blocks[0]
    xx = torch.from_numpy(z["x"])
    yy = blk(xx)
    x_rec, dx = blk.backward_pass(yy, torch.ones_like(yy))
    assert (x_rec - xx).abs().max().item() <= 1e-5 and dx.shape == xx.shape


def test_reversible_context_tensor_gets_summed_gradient():
    """A tensor the blocks close over (the encoder output in the decoder) receives the sum of every block's gradient once."""
    from reformer_tts_b200.model.reversible import ReversibleHalfResidual, ReversibleSequence, ReversibleSwap

    class AddCtx(nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = nn.Linear(8, 8)

        def forward(self, x, key=None):
            return self.lin(x) * key.mean(dim=1, keepdim=True)
    torch.manual_seed(0)
    blocks = nn.ModuleList([ReversibleHalfResidual(AddCtx()), ReversibleSwap(), ReversibleHalfResidual(AddCtx()), ReversibleSwap()])
    seq = ReversibleSequence(blocks)
    x = torch.randn(2, 4, 16, requires_grad=True)
    key = torch.randn(2, 3, 8, requires_grad=True)
    kw = [{"key": key}, {}, {"key": key}, {}]
    seq(x, kwargs_list=kw).sum().backward()
    gx, gk = x.grad.clone(), key.grad.clone()
    x.grad = key.grad = None
    h1, h2 = x.chunk(2, dim=2)
    h1 = h1 + blocks[0].f.net(h2, key=key)
    h1, h2 = h2, h1
    h1 = h1 + blocks[2].f.net(h2, key=key)
    (h1.sum() + h2.sum()).backward()
    assert (gx - x.grad).abs().max().item() <= 1e-5 and (gk - key.grad).abs().max().item() <= 1e-5


def test_reversible_blocks_offer_residual_and_gradient_accumulation():
    """Host protocol of the fused residual stream (reformer_tts_b200/residual.py): a sub-network that advertises
    ``takes_residual`` is offered the residual of the forward (``x + f``), of the reconstruction (``y - f``) and the gradient it is
    accumulated onto (``dx += df``); one that takes the offers must give the same outputs and gradients as one that ignores them,
    and a sub-network without the flag is never offered anything.  CPU stand-in for the GEMM / LayerNorm-backward epilogues."""
    from reformer_tts_b200.model.reversible import ReversibleBlock, ReversibleHalfResidual, ReversibleSequence, ReversibleSwap
    from reformer_tts_b200.residual import GradAccumRequest, ResidualRequest
    seen = {"add": 0, "reconstruct": 0, "acc": 0}

    class _TanhLinearFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w, fuse):
            y = torch.tanh(x @ w.t())
            ctx.save_for_backward(x, w, y)
            ctx.fuse = fuse
            resid, sub = ResidualRequest.take(x.shape, x.device) if fuse else (None, False)
            if resid is None:
                return y
            seen["reconstruct" if sub else "add"] += 1
            return (resid.view(x.shape) - y) if sub else (resid.view(x.shape) + y)

        @staticmethod
        def backward(ctx, dy):       # the incoming gradient is d loss / d f in every mode (see ResidualRequest)
            x, w, y = ctx.saved_tensors
            dz = dy * (1 - y * y)
            dx = dz @ w
            base = GradAccumRequest.take(dx.numel(), dx.device) if ctx.fuse else None
            if base is not None:
                seen["acc"] += 1
                dx = base.view(dx.shape) + dx
            return dx, dz.reshape(-1, dz.shape[-1]).t() @ x.reshape(-1, x.shape[-1]), None

    class Layer(nn.Module):
        def __init__(self, fuse, flag):
            super().__init__()
            self.w = nn.Parameter(torch.randn(8, 8) * 0.3)
            self.fuse = fuse
            if flag:
                self.takes_residual = True

        def forward(self, x):
            return _TanhLinearFn.apply(x, self.w, self.fuse)

    def run(fuse, flag):
        torch.manual_seed(0)
        nets = [Layer(fuse, flag) for _ in range(4)]
        blocks = nn.ModuleList([ReversibleBlock(nets[0], nets[1]), ReversibleHalfResidual(nets[2]), ReversibleSwap(),
                                ReversibleHalfResidual(nets[3]), ReversibleSwap()])
        seq = ReversibleSequence(blocks).train()
        torch.manual_seed(1)
        x = torch.randn(2, 5, 16, requires_grad=True)
        y = seq(x, kwargs_list=[{}] * 5)
        (y * torch.linspace(-1, 1, y.numel()).view_as(y)).sum().backward()
        return y.detach(), x.grad, [n.w.grad for n in nets]

    y0, gx0, gw0 = run(fuse=False, flag=False)
    assert seen == {"add": 0, "reconstruct": 0, "acc": 0}
    y1, gx1, gw1 = run(fuse=True, flag=True)
    # 4 sub-networks: every forward, reconstruction and gradient accumulation went through the requests (the first block's
    # residuals are halves of one tensor, i.e. not contiguous: those two offers are declined and added by the block)
    assert seen["reconstruct"] == 4 and seen["acc"] == 4 and 2 <= seen["add"] <= 4
    assert (y0 - y1).abs().max().item() <= 1e-6 and (gx0 - gx1).abs().max().item() <= 1e-6
    for a, b in zip(gw0, gw1):
        assert (a - b).abs().max().item() <= 1e-6
    before = dict(seen)
    y2, gx2, _ = run(fuse=True, flag=False)          # would take offers, but does not advertise it: nothing is offered
    assert seen == before and (y0 - y2).abs().max().item() <= 1e-6 and (gx0 - gx2).abs().max().item() <= 1e-6


def test_optimizer_groups_and_warmup_follow_the_reference_rules():
    """SURVEY.md 8(f) rank 2 (host logic only): AdamW parameter groups keyed on the substrings "bias" / "norm.weight"
    (ref:reformer_tts/training/wrappers.py:240-250) and the linear learning-rate warm-up (ref:...wrappers.py:286-295)."""
    from reformer_tts_b200.model import ReformerTTS, config as C
    from reformer_tts_b200.training import make_optimizer, param_groups, set_lr, warmup_lr
    model = ReformerTTS(**C.reference_model_kwargs("bucket-size-64-18-06"))
    groups = param_groups(model, 1e-6)
    names = dict(model.named_parameters())
    want_no_decay = {n for n in names if "bias" in n or "norm.weight" in n}
    got_no_decay = {n for n, p in names.items() if any(p is q for q in groups[1]["params"])}
    assert got_no_decay == want_no_decay and groups[1]["weight_decay"] == 0.0 and groups[0]["weight_decay"] == 1e-6
    assert len(groups[0]["params"]) + len(groups[1]["params"]) == len(names)
    # every LayerNorm gain and every bias is exempt; projection / embedding weights are not
    assert any(n.endswith("norm.weight") for n in got_no_decay) and all(not n.endswith("toqk.weight") for n in got_no_decay)
    opt = make_optimizer(model, 3e-4, 1e-6, fused=False)
    assert [g["weight_decay"] for g in opt.param_groups] == [1e-6, 0.0] and all(g["lr"] == 3e-4 for g in opt.param_groups)
    # warm-up of config/bucket-size-64-18-06.yml: 320 steps to 3e-4
    assert warmup_lr(0, 3e-4, 320) == pytest.approx(3e-4 / 320) and warmup_lr(159, 3e-4, 320) == pytest.approx(1.5e-4)
    assert warmup_lr(319, 3e-4, 320) == pytest.approx(3e-4) and warmup_lr(320, 3e-4, 320) == 3e-4 and warmup_lr(5, 3e-4, None) == 3e-4
    set_lr(opt, warmup_lr(0, 3e-4, 320))
    assert all(g["lr"] == pytest.approx(3e-4 / 320) for g in opt.param_groups)


def test_chunk_and_withnorm_fuse_only_rowwise_functions():
    from reformer_tts_b200.model import Chunk, WithNorm
    calls = []

    class Rowwise(nn.Module):
        rowwise = True

        def forward(self, x):
            calls.append(x.shape[-2])
            return x * 2

    class NotRowwise(nn.Module):
        def forward(self, x):
            calls.append(x.shape[-2])
            return x * 2
    x = torch.randn(2, 256, 8)
    assert torch.equal(Chunk(100, Rowwise(), along_dim=-2)(x), x * 2) and calls == [256]
    calls.clear()
    assert torch.equal(Chunk(100, NotRowwise(), along_dim=-2)(x), x * 2) and len(calls) == 86     # torch.chunk(256, 100) -> 86 pieces
    wn = WithNorm(nn.LayerNorm, 8, NotRowwise())
    assert not wn.rowwise and WithNorm(nn.LayerNorm, 8, Rowwise()).rowwise
    assert torch.allclose(wn(x), nn.functional.layer_norm(x, (8,)) * 2, atol=1e-6)


def test_shard_batch_partitions_exactly():
    from reformer_tts_b200.distributed import shard_batch
    for total, world in [(64, 8), (20, 8), (7, 2), (3, 4)]:
        spans = [shard_batch(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from torch import nn
from reformer_tts_b200.distributed import GradientAverager, shard_batch
from reformer_tts_b200.model.reversible import ReversibleBlock, ReversibleSequence
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
torch.manual_seed(0)
def mlp(): return nn.Sequential(nn.LayerNorm(8), nn.Linear(8, 16), nn.ReLU(), nn.Linear(16, 8))
class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.inp = nn.Linear(8, 8)
        self.layers = ReversibleSequence(nn.ModuleList([ReversibleBlock(mlp(), mlp()) for _ in range(3)]))
    def forward(self, x):
        h = self.inp(x)
        return self.layers(torch.cat([h, h], -1), kwargs_list=[{}] * 3).sum()
net = Net()
data = torch.randn(6, 5, 8)
lo, hi = shard_batch(6, dist.get_rank(), 2)
avg = GradientAverager(net, overlap=(sys.argv[4] == "1"))
assert avg.buckets.attached() and len(avg.buckets.groups) == 4       # three reversible blocks + one bucket per module outside the stack (inp)
# micro-batch 1 of 2: local accumulation only (no collective), micro-batch 2: reduce
avg.sync_enabled = False
(0.5 * net(data[lo:hi]) / (hi - lo)).backward()
avg.finish()
local = avg.buckets.flat.clone()
avg.sync_enabled = True
(0.5 * net(data[lo:hi]) / (hi - lo)).backward()
avg.finish()
assert avg.buckets.attached()
if dist.get_rank() == 0:
    ref = Net(); ref.load_state_dict(net.state_dict())
    (0.5 * (ref(data[:3]) / 3 + ref(data[3:]) / 3)).backward()
    err = max((p.grad - q.grad).abs().max().item() for p, q in zip(net.parameters(), ref.parameters()))
    half = Net(); half.load_state_dict(net.state_dict())
    (0.5 * half(data[:3]) / 3).backward()
    err_local = max((a - q.grad.reshape(-1)).abs().max().item() for a, q in
                    zip([local[o:o + p.numel()] for o, p in [(p.grad.storage_offset(), p) for p in net.parameters()]], half.parameters()))
    print("MAXERR", max(err, err_local))
dist.destroy_process_group()
"""


@pytest.mark.parametrize("overlap", ["0", "1"])
def test_data_parallel_gradient_average_gloo_world2(tmp_path, overlap):
    """World size 2 on CPU (gloo): flat gradient buckets (one per reversible block + the rest), per-bucket all-reduce from the
    reversible backward's block hook (overlap) or one flat all-reduce, and a non-boundary micro-batch that must stay local."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + (os.getpid() % 400) + (50 if overlap == "1" else 0))
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r), overlap], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=240) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    err = float(re.search(r"MAXERR ([0-9.e+-]+)", outs[0][0]).group(1))
    assert err <= 1e-5, err


# ---------------------------------------------------------------------------------------------- trainer rules (host logic, CPU)
class _ToyTTS(nn.Module):
    """Smallest module with the ReformerTTS call signature (phonemes, spectrogram, mask) -> (raw, post, stop, attention): a
    reversible stack between two linear layers, so TrainStep's bucket / accumulation / clipping logic runs on the CPU."""

    def __init__(self):
        super().__init__()
        from reformer_tts_b200.model.reversible import ReversibleBlock, ReversibleSequence
        mlp = lambda: nn.Sequential(nn.LayerNorm(16), nn.Linear(16, 32), nn.ReLU(), nn.Linear(32, 16))
        self.inp = nn.Linear(80, 16)
        self.layers = ReversibleSequence(nn.ModuleList([ReversibleBlock(mlp(), mlp()) for _ in range(2)]))
        self.mel = nn.Linear(16, 80)
        self.stop = nn.Linear(16, 1)

    def forward(self, phonemes, spectrogram, mask):
        h = self.inp(spectrogram)
        y = self.layers(torch.cat([h, h], -1), kwargs_list=[{}] * 2)
        h = torch.stack(y.chunk(2, dim=-1)).sum(0)
        mel = self.mel(h)
        return mel, mel + 0.1 * torch.tanh(mel), self.stop(h), None


def _toy_batch(n, seed):
    g = torch.Generator().manual_seed(seed)
    frames = 12
    spec = torch.randn(n, frames + 1, 80, generator=g) * 3
    stop = torch.zeros(n, frames)
    stop[:, -1] = 1
    return {"phonemes": torch.ones(n, 4, dtype=torch.long), "spectrogram": spec, "stop_tokens": stop, "loss_mask": torch.ones(n, frames, 80)}


def test_train_step_accumulation_clipping_and_warmup_follow_the_reference_rules():
    """ref:reformer_tts/training/train.py:77-89 (Lightning: accumulate_grad_batches, gradient_clip_val) and
    ref:reformer_tts/training/wrappers.py:284-297 (warm-up) against a plain-PyTorch statement of the same rules:
    k micro-batches of loss / k == one batch k times the size; clip_grad_norm_(1.0) before the update; lr set per optimiser step."""
    import copy
    from reformer_tts_b200.model.loss import TTSLoss
    from reformer_tts_b200.training import TrainStep, make_optimizer, set_lr, warmup_lr
    torch.manual_seed(0)
    base = _ToyTTS()
    loss_fn = TTSLoss(torch.tensor(5.))
    micro = [_toy_batch(2, s) for s in (1, 2, 3)]
    big = {k: torch.cat([m[k] for m in micro]) for k in micro[0]}

    # reference rules, stated with stock torch
    ref = copy.deepcopy(base)
    opt_r = make_optimizer(ref, 1e-2, 1e-6, fused=False)
    set_lr(opt_r, warmup_lr(0, 1e-2, 4))
    spec = big["spectrogram"]
    raw, post, stop, _ = ref(big["phonemes"], spec[:, :-1], big["loss_mask"].mean(-1))
    loss_r = loss_fn(raw, post, stop.view(6, -1), spec[:, 1:], big["stop_tokens"], big["loss_mask"])[0]
    loss_r.backward()
    norm_r = torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
    assert norm_r > 1.0, "the case must actually clip"
    opt_r.step()

    ours = copy.deepcopy(base)
    opt = make_optimizer(ours, 1e-2, 1e-6, fused=False)
    step = TrainStep(ours, loss_fn, opt, micro[0], use_cuda_graph=False, grad_clip=1.0, accumulate_grad_batches=3)
    set_lr(opt, warmup_lr(step.optimizer_steps, 1e-2, 4))
    before = [p.detach().clone() for p in ours.parameters()]
    losses = [step.step(micro[0]), step.step(micro[1])]
    assert step.optimizer_steps == 0 and all(torch.equal(a, b) for a, b in zip(before, ours.parameters())), "no update before the boundary"
    losses.append(step.step(micro[2]))
    assert step.optimizer_steps == 1
    assert torch.allclose(torch.stack(losses).mean(), loss_r, rtol=1e-5)
    assert torch.allclose(step.grad_norm, norm_r, rtol=1e-4)
    for a, b in zip(ours.parameters(), ref.parameters()):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6)
    assert float(step.buckets.flat.abs().max()) == 0.0 and step.buckets.attached()      # zeroed in place, still the parameters' .grad
    # a second optimiser step runs at the next warm-up rate
    set_lr(opt, warmup_lr(step.optimizer_steps, 1e-2, 4))
    assert opt.param_groups[0]["lr"] == pytest.approx(2 * 1e-2 / 4)


def test_train_step_refuses_to_capture_an_optimizer_whose_lr_would_be_baked_in():
    """a float lr (or a non-capturable optimiser) must not reach CUDA-graph capture, where warm-up would silently stop working."""
    from reformer_tts_b200.model.loss import TTSLoss
    from reformer_tts_b200.training import TrainStep, make_optimizer
    model = _ToyTTS()
    opt = make_optimizer(model, 1e-3, 1e-6, fused=False, capturable=False)
    with pytest.raises(ValueError, match="capturable"):
        TrainStep(model, TTSLoss(torch.tensor(5.)), opt, _toy_batch(2, 0), use_cuda_graph=True)


def test_gradient_buckets_layout():
    from reformer_tts_b200.distributed import GradientBuckets
    model = _ToyTTS()
    model.stop.weight.requires_grad_(False)
    b = GradientBuckets(model)
    # two reversible blocks, then one bucket per module outside the stack in parameter order: inp, mel, stop (bias only)
    assert len(b.groups) == 5 and b.rest_buckets == [2, 3, 4] and [len(g) for g in b.groups[2:]] == [2, 2, 1]
    assert all(start % 64 == 0 for start, _ in b.bounds) and b.flat.numel() % 64 == 0
    n_train = sum(p.numel() for p in model.parameters() if p.requires_grad)
    assert sum(p.numel() for g in b.groups for p in g) == n_train and model.stop.weight.grad is None
    for i, blk in enumerate(model.layers.blocks):
        assert b.bucket_of_block(blk) == i
        lo, hi = b.bounds[i]
        for p in blk.parameters():
            assert lo <= p.grad.storage_offset() and p.grad.storage_offset() + p.numel() <= hi
    model(torch.ones(1, 4, dtype=torch.long), torch.randn(1, 12, 80), None)[0].sum().backward()
    assert b.attached() and float(b.flat.abs().sum()) > 0
    b.zero()
    assert all(float(p.grad.abs().sum()) == 0 for g in b.groups for p in g)
