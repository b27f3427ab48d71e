"""Modules of ref:reformer_tts/model/modules.py.

``FeedForward`` is on the hot path and runs on the fused kernels.  The pre/post nets and the positional encoding
are OUTSIDE the hot-path scope: they are restated here on stock PyTorch layers,
with the reference's state-dict keys, only so that a complete ReformerTTS training step can be timed on the GPU box
(where /root/reference does not exist)."""
from __future__ import annotations

from collections import OrderedDict

import torch
from torch import nn

from ..feed_forward import ln_feed_forward
from ..lsh_attention import _WeightCache


class FeedForward(nn.Module):
    """Linear(dim, hidden) - ReLU - Dropout - Linear(hidden, dim); parameters at ``net.0`` / ``net.3``
    (ref:...modules.py:195-207)."""
    rowwise = True
    takes_residual = True      # its last GEMM can write x + f(.) / y - f(.) (reformer_tts_b200.residual.ResidualRequest)

    def __init__(self, dim=512, hidden=2048, dropout=0.):
        super().__init__()
        if dropout:
            raise NotImplementedError("FeedForward dropout > 0 is not built (0 in every reference config)")
        if dim % 128 or hidden % 128:
            raise NotImplementedError("FeedForward: dim and hidden must be multiples of 128")
        self.net = nn.Sequential(nn.Linear(dim, hidden), nn.ReLU(), nn.Dropout(dropout), nn.Linear(hidden, dim))
        self._c1, self._c2 = _WeightCache(), _WeightCache()

    def forward_with_norm(self, x, norm):
        return ln_feed_forward(x, norm, self.net[0], self.net[3], self._c1, self._c2)

    def forward(self, x):
        return self.forward_with_norm(x, None)


def _conv5(cin, cout):
    return nn.Conv1d(cin, cout, kernel_size=5, padding=2)


def _conv_stack_tokens_last(seq: nn.Sequential, x: torch.Tensor) -> torch.Tensor:
    """``seq(x.transpose(1, 2)).transpose(1, 2)`` for a stack of Conv1d / BatchNorm1d / elementwise layers, x = [B, T, C].

    On the GPU the token-major activation IS the channels-last image [B, C, 1, T] (strides T*C, 1, -, C), so the stack runs on 4-D
    channels-last views: the vendor convolutions take it as is (their implicit-GEMM kernels are NHWC; with the reference's [B, C, T]
    layout every convolution was wrapped in an NCHW->NHWC / NHWC->NCHW transpose pair - 56 launches, ~0.6 ms of a 22.7 ms step) and
    both transposes around the stack are views.  Same modules, same parameters, same state-dict keys, same arithmetic (BatchNorm1d's
    running-statistics bookkeeping included); the CPU path (oracle) keeps the reference's layout."""
    if not x.is_cuda:
        return seq(x.transpose(1, 2)).transpose(1, 2)
    import torch.nn.functional as F
    h = x.transpose(1, 2).unsqueeze(2)                      # [B, C, 1, T]: a view, contiguous in the channels-last sense
    for m in seq:
        if isinstance(m, nn.Conv1d):
            w = m.weight.unsqueeze(2).contiguous(memory_format=torch.channels_last)
            h = F.conv2d(h, w, m.bias, stride=(1, m.stride[0]), padding=(0, m.padding[0]), dilation=(1, m.dilation[0]), groups=m.groups)
        elif isinstance(m, nn.BatchNorm1d):
            factor = 0.0 if m.momentum is None else m.momentum
            if m.training and m.track_running_stats and m.num_batches_tracked is not None:
                m.num_batches_tracked.add_(1)
                if m.momentum is None:
                    factor = 1.0 / float(m.num_batches_tracked)
            use_batch_stats = m.training or (m.running_mean is None and m.running_var is None)
            h = F.batch_norm(h, m.running_mean if not m.training or m.track_running_stats else None,
                             m.running_var if not m.training or m.track_running_stats else None, m.weight, m.bias, use_batch_stats, factor, m.eps)
        else:
            h = m(h)
    return h.squeeze(2).transpose(1, 2)


class EncoderPreNet(nn.Module):
    """Embedding -> 3 x (dropout, conv5, batch-norm, ReLU) -> dropout -> Linear  (ref:...modules.py:8-61)."""

    def __init__(self, num_embeddings: int, embedding_dim: int = 512, dropout: float = 0.5):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.embed = nn.Embedding(num_embeddings, embedding_dim, padding_idx=0)
        self.projection = nn.Linear(embedding_dim, embedding_dim)
        layers = [("dropout0", nn.Dropout(dropout))]
        for i in (1, 2, 3):
            layers += [(f"conv{i}", _conv5(embedding_dim, embedding_dim)), (f"bn{i}", nn.BatchNorm1d(embedding_dim)),
                       (f"relu{i}", nn.ReLU()), (f"dropout{i}", nn.Dropout(dropout))]
        self.convolutions = nn.Sequential(OrderedDict(layers))

    def forward(self, tokens):
        return self.projection(_conv_stack_tokens_last(self.convolutions, self.embed(tokens)))


class DecoderPreNet(nn.Module):
    """Linear-ReLU-dropout x2 then a projection  (ref:...modules.py:64-100)."""

    def __init__(self, input_size: int, output_size: int, hidden_size: int = 256, dropout: float = 0.5):
        super().__init__()
        self.input_size, self.output_size, self.hidden_size = input_size, output_size, hidden_size
        self.layer = nn.Sequential(OrderedDict([
            ("fc1", nn.Linear(input_size, hidden_size)), ("relu1", nn.ReLU()), ("dropout1", nn.Dropout(dropout)),
            ("fc2", nn.Linear(hidden_size, output_size)), ("relu2", nn.ReLU()), ("dropout2", nn.Dropout(dropout)),
            ("projection", nn.Linear(output_size, output_size)),
        ]))

    def forward(self, mel):
        return self.layer(mel)


class PostConvNet(nn.Module):
    """``depth`` x (conv5, batch-norm, tanh, dropout) + conv5 back to mel  (ref:...modules.py:103-169)."""

    def __init__(self, mel_size: int, num_hidden: int, dropout: float, depth: int):
        super().__init__()
        self.mel_size = mel_size
        layers = []
        for i in range(depth):
            layers += [(f"conv{i}", _conv5(mel_size if i == 0 else num_hidden, num_hidden)), (f"bn{i}", nn.BatchNorm1d(num_hidden)),
                       (f"tanh{i}", nn.Tanh()), (f"dropout{i}", nn.Dropout(dropout))]
        layers.append(("convend", _conv5(num_hidden, mel_size)))
        self.layers = nn.Sequential(OrderedDict(layers))

    def forward(self, mel):
        return _conv_stack_tokens_last(self.layers, mel)


class ScaledPositionalEncoding(nn.Module):
    """x + alpha * dropout(interleaved sin/cos table)  (ref:...modules.py:172-192)."""

    def __init__(self, d_model, dropout):
        super().__init__()
        self.register_buffer("inv_freq", 1. / (10000 ** (torch.arange(0, d_model, 2).float() / d_model)))
        self.alpha = nn.Parameter(torch.empty(1).normal_(0, 1))
        self.dropout = nn.Dropout(dropout)

    def forward(self, x):
        pos = torch.arange(x.shape[1], device=x.device, dtype=self.inv_freq.dtype)
        angle = pos[:, None] * self.inv_freq[None, :]
        table = torch.stack([angle.sin(), angle.cos()], dim=-1).flatten(1)      # [T, d_model], sin/cos interleaved
        return x + self.alpha * self.dropout(table)
