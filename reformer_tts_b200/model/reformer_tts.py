"""The caller of the hot path: the ReformerTTS model of ref:reformer_tts/model/reformer_tts.py (training ``forward``
only; the autoregressive ``infer`` loop is out of scope).  Same constructor kwargs and
state-dict keys (``enc.prenet.*``, ``enc.reformer.layers.blocks.N.f.net...``, ``dec.*``, ``postnet.*``).  Device-aware
padding (the reference builds its pad tensors on the default device)."""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import nn

from .modules import DecoderPreNet, EncoderPreNet, PostConvNet, ScaledPositionalEncoding
from .reformer import ReformerDec, ReformerEnc


def pad_to_multiple(tensor: torch.Tensor, pad_base: int) -> torch.Tensor:
    """Zero-pad dim 1 of a (batch, length, channels) tensor up to a multiple of pad_base (ref:...reformer_tts.py:224-232)."""
    length = tensor.shape[1]
    target = ((length - 1) // pad_base + 1) * pad_base
    if target == length:
        return tensor
    return torch.nn.functional.pad(tensor, (0, 0, 0, target - length))


class Encoder(nn.Module):
    def __init__(self, dict_size: int, embedding_dim: int, scp_encoding_dropout: float, reformer_kwargs: Dict, prenet_kwargs: Dict):
        super().__init__()
        self.prenet = EncoderPreNet(num_embeddings=dict_size + 1, embedding_dim=embedding_dim, **prenet_kwargs)
        self.positional_encoding = ScaledPositionalEncoding(embedding_dim, scp_encoding_dropout)
        self.reformer = ReformerEnc(embedding_dim, **reformer_kwargs)

    def forward(self, tokens, input_mask=None):
        return self.reformer(self.positional_encoding(self.prenet(tokens)), input_mask=input_mask)


class Decoder(nn.Module):
    def __init__(self, num_mel_coeffs: int, embedding_dim: int, scp_encoding_dropout: float, prenet_kwargs: Dict, reformer_kwargs: Dict):
        super().__init__()
        self.prenet = DecoderPreNet(input_size=num_mel_coeffs, output_size=embedding_dim, **prenet_kwargs)
        self.positional_encoding = ScaledPositionalEncoding(embedding_dim, scp_encoding_dropout)
        self.reformer = ReformerDec(embedding_dim, **reformer_kwargs)
        self.mel_linear = nn.Linear(embedding_dim, num_mel_coeffs)
        self.stop_linear = nn.Linear(embedding_dim, 1)

    def forward(self, mel, keys, key_padding_mask=None, input_mask=None):
        hidden, attention_matrices = self.reformer(self.positional_encoding(self.prenet(mel)), keys=keys,
                                                   key_padding_mask=key_padding_mask, input_mask=input_mask)
        return self.mel_linear(hidden), self.stop_linear(hidden), attention_matrices


class ReformerTTS(nn.Module):
    def __init__(self, num_mel_coeffs: int, dict_size: int, pad_base: int, embedding_dim: int, scp_encoding_dropout: float,
                 enc_reformer_kwargs: Dict, enc_prenet_kwargs: Dict, dec_prenet_kwargs: Dict, dec_reformer_kwargs: Dict,
                 postnet_kwargs: Dict):
        super().__init__()
        self.num_mel_coeffs = num_mel_coeffs
        self.pad_base = pad_base
        self.enc = Encoder(dict_size=dict_size, embedding_dim=embedding_dim, scp_encoding_dropout=scp_encoding_dropout,
                           reformer_kwargs=enc_reformer_kwargs, prenet_kwargs=enc_prenet_kwargs)
        self.dec = Decoder(num_mel_coeffs=num_mel_coeffs, embedding_dim=embedding_dim, scp_encoding_dropout=scp_encoding_dropout,
                           prenet_kwargs=dec_prenet_kwargs, reformer_kwargs=dec_reformer_kwargs)
        self.postnet = PostConvNet(mel_size=num_mel_coeffs, num_hidden=embedding_dim, **postnet_kwargs)

    def forward(self, phonemes: torch.Tensor, spectrogram: torch.Tensor, spectrogram_mask: Optional[torch.Tensor] = None):
        """phonemes (B, Lp) int; spectrogram (B, Lm, n_mels); mask (B, Lm), non-zero = real frame.
        Returns (mel, mel + postnet residual, stop logits, attention matrices), cut back to Lm frames
        (ref:...reformer_tts.py:103-143)."""
        dev = spectrogram.device
        phonemes = phonemes.to(dev)
        pad_phonemes = pad_to_multiple(phonemes.unsqueeze(-1), self.pad_base).squeeze(-1)
        phoneme_mask = pad_phonemes != 0
        if spectrogram_mask is None:
            spectrogram_mask = torch.ones(spectrogram.shape[:2], device=dev)
        frame_mask = pad_to_multiple(spectrogram_mask.to(dev).unsqueeze(-1), self.pad_base).squeeze(-1).to(torch.bool)
        keys = self.enc(pad_phonemes, input_mask=phoneme_mask)
        mel, stop, attention_matrices = self.dec(pad_to_multiple(spectrogram, self.pad_base), keys=keys,
                                                 key_padding_mask=~phoneme_mask, input_mask=frame_mask)
        mel_postnet = mel + self.postnet(mel)
        cut = spectrogram.shape[1]
        return mel[:, :cut], mel_postnet[:, :cut], stop[:, :cut], attention_matrices
