"""CPU restatement of the reference model around the hot path (TEST INFRASTRUCTURE ONLY): ReformerEnc / ReformerDec
(ref:reformer_tts/model/reformer.py:51-158) with real chunking, separate LayerNorm, the oracle LSH layers and the
oracle reversible machinery, plus the full ReformerTTS assembled from them.  State-dict keys equal the product's, so
one set of weights loads into both.  Used by the parity tests and as bench.py's CPU baseline / ``--impl reference`` arm
(the reference itself cannot travel to the GPU box and its reformer_pytorch dependency is not installable).

The pre/post nets, positional encoding and loss are stock-PyTorch modules shared with the product package: they are
outside the hot path and contain no CUDA-extension call."""
from __future__ import annotations

from typing import Dict

import torch
from torch import nn

from reformer_tts_b200.model.modules import DecoderPreNet, EncoderPreNet, PostConvNet, ScaledPositionalEncoding
from reformer_tts_b200.model.reformer_tts import pad_to_multiple

from .lsh_hf import LSHSelfAttentionHF
from .lsh_rp import LSHSelfAttentionRP
from .reversible import RevBlock, RevHalf, RevSequence, RevSwap


class FeedForward(nn.Module):       # ref:reformer_tts/model/modules.py:195-207
    def __init__(self, dim=512, hidden=2048, dropout=0.):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden), nn.ReLU(), nn.Dropout(dropout), nn.Linear(hidden, dim))
        self.round_operands = False      # oracle/rounded.py

    def forward(self, x):
        if self.round_operands:
            from .rounded import FeedForwardFn
            assert self.net[2].p == 0 or not self.training, "rounding mode: hidden dropout is 0 in every reference config"
            return FeedForwardFn.apply(x, self.net[0].weight, self.net[0].bias, self.net[3].weight, self.net[3].bias)
        return self.net(x)


class WithNorm(nn.Module):          # ref:reformer_tts/model/reformer.py:25-33
    def __init__(self, norm_class, dim, fn):
        super().__init__()
        self.norm, self.fn = norm_class(dim), fn

    def forward(self, x, **kwargs):
        return self.fn(self.norm(x), **kwargs)


class Chunk(nn.Module):             # ref:reformer_tts/model/reformer.py:36-45
    def __init__(self, chunks, fn, along_dim=-1):
        super().__init__()
        self.dim, self.chunks, self.fn = along_dim, chunks, fn

    def forward(self, x):
        return torch.cat([self.fn(c) for c in x.chunk(self.chunks, dim=self.dim)], dim=self.dim)


class LSHWrapper(nn.Module):        # ref:reformer_tts/model/reformer.py:189-220
    def __init__(self, dim, causal, **kwargs):
        super().__init__()
        kwargs = dict(kwargs)
        self.implementation = kwargs.pop("implementation")
        if self.implementation == "reformer_pytorch":
            self.layer = LSHSelfAttentionRP(dim, causal=causal, **kwargs)
        else:
            self.layer = LSHSelfAttentionHF(dim, kwargs["heads"], kwargs["bucket_size"], kwargs["n_hashes"], causal, kwargs["dropout"])

    def forward(self, x, input_mask=None):
        if self.implementation == "reformer_pytorch":
            return self.layer(x, input_mask=input_mask)
        return self.layer(x, attention_mask=input_mask)


class CrossAttention(nn.Module):    # ref:reformer_tts/model/reformer.py:161-186
    def __init__(self, dim, attention_matrices=None, **kwargs):
        super().__init__()
        self.layer = nn.MultiheadAttention(dim, **kwargs)
        self.attention_matrices_ = attention_matrices
        self.round_operands = False      # oracle/rounded.py

    def forward(self, query, **kwargs):
        if self.round_operands and self.training:
            from .rounded import CrossAttentionFn
            assert self.layer.dropout == 0, "rounding mode: attention dropout is 0 in every reference config"
            return CrossAttentionFn.apply(query, kwargs["key"], self.layer.in_proj_weight, self.layer.in_proj_bias, self.layer.out_proj.weight,
                                          self.layer.out_proj.bias, kwargs.get("key_padding_mask"), self.layer.num_heads)
        mem = kwargs["key"].transpose(0, 1)
        extra = {k: v for k, v in kwargs.items() if k not in ("key", "value")}
        out, w = self.layer(query.transpose(0, 1), mem, mem, **extra)
        if not self.training and self.attention_matrices_ is not None:
            self.attention_matrices_.append(w)
        return out.transpose(0, 1)


def _ff(dim, ff_chunks, ff_kwargs):
    ff = WithNorm(nn.LayerNorm, dim, FeedForward(dim, **ff_kwargs))
    return Chunk(ff_chunks, ff, along_dim=-2) if ff_chunks > 1 else ff


class ReformerEnc(nn.Module):
    def __init__(self, dim, depth, ff_chunks, attn_kwargs, ff_kwargs):
        super().__init__()
        self.depth = depth
        self.layers = RevSequence(nn.ModuleList([
            RevBlock(WithNorm(nn.LayerNorm, dim, LSHWrapper(dim, causal=False, **attn_kwargs)), _ff(dim, ff_chunks, ff_kwargs))
            for _ in range(depth)]))

    def forward(self, x, input_mask=None):
        kwargs_list = [{"f_args": {"input_mask": input_mask}} for _ in range(self.depth)]
        y = self.layers(torch.cat([x, x], dim=-1), kwargs_list=kwargs_list)
        return torch.stack(y.chunk(2, dim=-1)).sum(dim=0)


class ReformerDec(nn.Module):
    def __init__(self, dim, depth, ff_chunks, attn_kwargs, self_attn_kwargs, ff_kwargs):
        super().__init__()
        self.depth = depth
        self.attention_matrices_ = []
        blocks = []
        for _ in range(depth):
            blocks += [RevHalf(WithNorm(nn.LayerNorm, dim, LSHWrapper(dim, causal=True, **self_attn_kwargs))), RevSwap(),
                       RevHalf(WithNorm(nn.LayerNorm, dim, CrossAttention(dim, self.attention_matrices_, **attn_kwargs))), RevSwap(),
                       RevHalf(_ff(dim, ff_chunks, ff_kwargs)), RevSwap()]
        self.layers = RevSequence(nn.ModuleList(blocks))

    def forward(self, x, keys, key_padding_mask=None, input_mask=None):
        kwargs_list = [dict() for _ in range(6 * self.depth)]
        for kw in kwargs_list[2::6]:
            kw.update(key=keys, value=keys, key_padding_mask=key_padding_mask)
        for kw in kwargs_list[::6]:
            kw["input_mask"] = input_mask
        self.attention_matrices_.clear()
        y = self.layers(torch.cat([x, x], dim=-1), kwargs_list=kwargs_list)
        return torch.stack(y.chunk(2, dim=-1)).sum(dim=0), self.attention_matrices_


class _Enc(nn.Module):
    def __init__(self, dict_size, embedding_dim, scp_encoding_dropout, reformer_kwargs, prenet_kwargs):
        super().__init__()
        self.prenet = EncoderPreNet(num_embeddings=dict_size + 1, embedding_dim=embedding_dim, **prenet_kwargs)
        self.positional_encoding = ScaledPositionalEncoding(embedding_dim, scp_encoding_dropout)
        self.reformer = ReformerEnc(embedding_dim, **reformer_kwargs)

    def forward(self, tokens, input_mask=None):
        return self.reformer(self.positional_encoding(self.prenet(tokens)), input_mask=input_mask)


class _Dec(nn.Module):
    def __init__(self, num_mel_coeffs, embedding_dim, scp_encoding_dropout, prenet_kwargs, reformer_kwargs):
        super().__init__()
        self.prenet = DecoderPreNet(input_size=num_mel_coeffs, output_size=embedding_dim, **prenet_kwargs)
        self.positional_encoding = ScaledPositionalEncoding(embedding_dim, scp_encoding_dropout)
        self.reformer = ReformerDec(embedding_dim, **reformer_kwargs)
        self.mel_linear = nn.Linear(embedding_dim, num_mel_coeffs)
        self.stop_linear = nn.Linear(embedding_dim, 1)

    def forward(self, mel, keys, key_padding_mask=None, input_mask=None):
        hid, att = self.reformer(self.positional_encoding(self.prenet(mel)), keys=keys, key_padding_mask=key_padding_mask,
                                 input_mask=input_mask)
        return self.mel_linear(hid), self.stop_linear(hid), att


class ReformerTTSOracle(nn.Module):     # ref:reformer_tts/model/reformer_tts.py:70-143
    def __init__(self, num_mel_coeffs, dict_size, pad_base, embedding_dim, scp_encoding_dropout, enc_reformer_kwargs: Dict,
                 enc_prenet_kwargs: Dict, dec_prenet_kwargs: Dict, dec_reformer_kwargs: Dict, postnet_kwargs: Dict):
        super().__init__()
        self.pad_base = pad_base
        self.enc = _Enc(dict_size, embedding_dim, scp_encoding_dropout, enc_reformer_kwargs, enc_prenet_kwargs)
        self.dec = _Dec(num_mel_coeffs, embedding_dim, scp_encoding_dropout, dec_prenet_kwargs, dec_reformer_kwargs)
        self.postnet = PostConvNet(mel_size=num_mel_coeffs, num_hidden=embedding_dim, **postnet_kwargs)

    def forward(self, phonemes, spectrogram, spectrogram_mask=None):
        pad_ph = pad_to_multiple(phonemes.unsqueeze(-1), self.pad_base).squeeze(-1)
        ph_mask = pad_ph != 0
        if spectrogram_mask is None:
            spectrogram_mask = torch.ones(spectrogram.shape[:2])
        fr_mask = pad_to_multiple(spectrogram_mask.unsqueeze(-1), self.pad_base).squeeze(-1).to(torch.bool)
        keys = self.enc(pad_ph, input_mask=ph_mask)
        mel, stop, att = self.dec(pad_to_multiple(spectrogram, self.pad_base), keys=keys, key_padding_mask=~ph_mask, input_mask=fr_mask)
        post = mel + self.postnet(mel)
        cut = spectrogram.shape[1]
        return mel[:, :cut], post[:, :cut], stop[:, :cut], att
