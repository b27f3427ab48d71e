"""Generates the golden vectors under tests/golden/ from the REAL implementations available in the build container:

* hf_lsh_*.npz      - transformers (5.5.0) ``LSHSelfAttention`` configured exactly as
                      ref:reformer_tts/model/reformer.py:204-213, run through the reference's own
                      ``LSHSelfAttentionWrapper`` (imported from /root/reference with the 2.11-era module path aliased);
                      inputs, weights, rotations seed, bucket ids and hidden states.
* hfgrad_*.npz      - the same class, forward AND backward (d/dx, d/dWqk, d/dWv), including the shapes of
                      config/huggingface-lsh.yml (dim 512, 8 heads, 8 rounds; T=256 chunk 64 / T=1024 chunk 128); inputs are
                      regenerated from the seed by the tests (``hf_inputs``), a checksum guards against generator drift.
* reversible_ref.npz - gradients produced by the reference's own ref:reformer_tts/model/reversible.py
                      (ReversibleSequence of ReversibleBlock / ReversibleHalfResidual / ReversibleSwap) on small MLP
                      sub-networks, plus the plain-autograd gradients through its IrreversibleBlock.
* state_dict_keys.json - parameter names/shapes of the reference ReformerTTS (HF variant; the RP variant cannot be built
                      because reformer_pytorch is not installable) for the checkpoint-compatibility test.

Run once here (needs /root/reference and transformers); the outputs are committed, the GPU box only reads them."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(HERE))
from _util import hf_inputs  # noqa: E402  (tests/_util.py: the tests regenerate the same inputs from the seed)
import transformers  # noqa: E402
import transformers.models.reformer.modeling_reformer as _mr  # noqa: E402

sys.modules["transformers.modeling_reformer"] = _mr     # import path of transformers 2.11 used at ref:...reformer.py:203
transformers.logging.set_verbosity_error()
from reformer_tts.model import reformer as ref_reformer  # noqa: E402
from reformer_tts.model import reversible as ref_rev  # noqa: E402


def hf_case(name, dim, heads, bucket, n_hashes, causal, T, B, pad, seed):
    torch.manual_seed(seed)
    wrapper = ref_reformer.LSHSelfAttentionWrapper(dim, causal, implementation="huggingface_transformers", heads=heads,
                                                   bucket_size=bucket, n_hashes=n_hashes, dropout=0.).eval()
    x = torch.randn(B, T, dim)
    mask = None
    if pad:
        mask = torch.ones(B, T, dtype=torch.bool)
        mask[0, T - 37:] = False
        mask[-1, T - 5:] = False
    torch.manual_seed(seed + 1)     # state right before the rotations are drawn
    out = wrapper.layer(x, attention_mask=mask)
    np.savez_compressed(os.path.join(HERE, f"hf_lsh_{name}.npz"),
                        x=x.numpy(), wqk=wrapper.layer.query_key.weight.detach().numpy(), wv=wrapper.layer.value.weight.detach().numpy(),
                        mask=np.zeros(0) if mask is None else mask.numpy(), hidden=out.hidden_states.detach().numpy(),
                        buckets=out.buckets.numpy().astype(np.int16), meta=np.array([dim, heads, bucket, n_hashes, int(causal), seed + 1]))
    print(name, "hidden", tuple(out.hidden_states.shape), "buckets", tuple(out.buckets.shape))


def hf_grad_case(name, dim, heads, bucket, n_hashes, causal, T, B, pad, seed, stride):
    """Forward AND backward of the real transformers class at a reference-config shape (config/huggingface-lsh.yml: dim 512,
    8 heads, 8 hash rounds; encoder chunk 64 at T=256, decoder chunk 128 at T=1024): hidden states, bucket ids, d/dx, d/dWqk,
    d/dWv.  Row-strided copies keep the fixture small (weight gradients sum over every token, so they pin all rows anyway)."""
    wrapper = ref_reformer.LSHSelfAttentionWrapper(dim, causal, implementation="huggingface_transformers", heads=heads,
                                                   bucket_size=bucket, n_hashes=n_hashes, dropout=0.).train()
    wqk, wv, x, dy, mask, checksum = hf_inputs(dim, T, B, pad, seed)
    with torch.no_grad():
        wrapper.layer.query_key.weight.copy_(wqk)
        wrapper.layer.value.weight.copy_(wv)
    x.requires_grad_(True)
    torch.manual_seed(seed + 1)     # state right before the rotations are drawn
    out = wrapper.layer(x, attention_mask=mask)
    out.hidden_states.backward(dy)
    np.savez_compressed(os.path.join(HERE, f"hfgrad_{name}.npz"),
                        hidden=out.hidden_states.detach().numpy()[:, ::stride], dx=x.grad.numpy()[:, ::stride],
                        dwqk=wrapper.layer.query_key.weight.grad.numpy()[::stride], dwv=wrapper.layer.value.weight.grad.numpy()[::stride],
                        buckets=out.buckets.numpy().astype(np.int16), checksum=np.array([checksum]),
                        meta=np.array([dim, heads, bucket, n_hashes, int(causal), T, B, int(pad), seed, stride]))
    print(name, "hidden", tuple(out.hidden_states.shape), "dx", tuple(x.grad.shape))


def mlp(d, seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.LayerNorm(d), torch.nn.Linear(d, 2 * d), torch.nn.ReLU(), torch.nn.Linear(2 * d, d))


def reversible_case():
    d, B, T = 16, 2, 8
    nets = [mlp(d, 10 + i) for i in range(7)]
    blocks = torch.nn.ModuleList([ref_rev.ReversibleBlock(nets[0], nets[1]), ref_rev.ReversibleBlock(nets[2], nets[3]),
                                  ref_rev.ReversibleHalfResidual(nets[4]), ref_rev.ReversibleSwap(),
                                  ref_rev.ReversibleHalfResidual(nets[5]), ref_rev.ReversibleSwap()])
    seq = ref_rev.ReversibleSequence(blocks).train()
    torch.manual_seed(0)
    x = torch.randn(B, T, 2 * d, requires_grad=True)
    w = torch.randn(B, T, 2 * d)
    y = seq(x, kwargs_list=[{}, {}, {}, {}, {}, {}])
    (y * w).sum().backward()
    out = {"x": x.detach().numpy(), "w": w.numpy(), "y": y.detach().numpy(), "dx": x.grad.numpy()}
    for i, n in enumerate(nets[:6]):
        for k, p in n.named_parameters():
            out[f"g{i}_{k}"] = p.grad.numpy()
    # plain autograd through the reference's IrreversibleBlock for the first two blocks (KAT-5)
    for n in nets:
        n.zero_grad()
    irr = [ref_rev.IrreversibleBlock(nets[0], nets[1]), ref_rev.IrreversibleBlock(nets[2], nets[3])]
    x2 = x.detach().clone().requires_grad_(True)
    h = x2
    for blk in irr:
        h = blk(h, {}, {})
    (h * w).sum().backward()
    out["irr_y"] = h.detach().numpy()
    out["irr_dx"] = x2.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "reversible_ref.npz"), **out)
    print("reversible", {k: v.shape for k, v in out.items() if k in ("y", "dx", "irr_dx")})


def state_dict_keys():
    from reformer_tts.model.reformer_tts import ReformerTTS
    attn = dict(implementation="huggingface_transformers", heads=8, bucket_size=64, n_hashes=8, add_local_attn_hash=False, attn_chunks=1,
                random_rotations_per_head=False, attend_across_buckets=True, allow_duplicate_attention=True, num_mem_kv=0,
                one_value_head=False, use_full_attn=False, full_attn_thres=None, return_attn=False, post_attn_dropout=0., dropout=0.)
    ff = dict(hidden=2048, dropout=0.)
    mha = dict(num_heads=8, dropout=0., bias=True, add_bias_kv=False, add_zero_attn=False, kdim=None, vdim=None)
    model = ReformerTTS(num_mel_coeffs=80, dict_size=76, pad_base=256, embedding_dim=512, scp_encoding_dropout=0.05,
                        enc_reformer_kwargs=dict(depth=3, ff_chunks=100, attn_kwargs=attn, ff_kwargs=ff), enc_prenet_kwargs=dict(dropout=0.05),
                        dec_prenet_kwargs=dict(hidden_size=256, dropout=0.05),
                        dec_reformer_kwargs=dict(depth=3, ff_chunks=100, attn_kwargs=mha, self_attn_kwargs=dict(attn, bucket_size=128), ff_kwargs=ff),
                        postnet_kwargs=dict(depth=2, dropout=0.1))
    keys = {k: list(v.shape) for k, v in model.state_dict().items()}
    with open(os.path.join(HERE, "state_dict_keys_hf.json"), "w") as fh:
        json.dump(keys, fh, indent=0, sort_keys=True)
    print("state dict keys", len(keys))


if __name__ == "__main__":
    hf_case("enc_full", 128, 2, 64, 4, False, 256, 2, False, 100)
    hf_case("enc_pad", 128, 2, 64, 4, False, 256, 2, True, 200)
    hf_case("dec_causal", 128, 2, 128, 2, True, 512, 2, False, 300)
    hf_case("dec_causal_pad", 128, 2, 128, 2, True, 512, 2, True, 400)
    hf_grad_case("small_dec", 128, 2, 64, 4, True, 256, 2, True, 500, 1)
    hf_grad_case("cfg3_enc", 512, 8, 64, 8, False, 256, 1, True, 600, 2)
    hf_grad_case("cfg3_dec", 512, 8, 128, 8, True, 1024, 1, True, 700, 4)
    reversible_case()
    state_dict_keys()
