"""TTSLoss of ref:reformer_tts/model/loss.py:7-53 (masked MSE/L1 on raw + postnet mel, BCE-with-logits on stop).
The masking is done out of place: the reference's in-place ``*=`` on a view that the postnet saved for backward is
rejected by current autograd."""
from __future__ import annotations

import torch
from torch import nn
from torch.nn.functional import binary_cross_entropy_with_logits


class TTSLoss(nn.Module):
    def __init__(self, pos_weight: torch.Tensor, raw_pred_loss_weight: float = 1., post_pred_loss_weight: float = 1.,
                 stop_loss_weight: float = 1., spectrogram_loss: str = "mse"):
        super().__init__()
        if spectrogram_loss not in ("mse", "l1"):
            raise RuntimeError(f"Unsupported loss type: {spectrogram_loss}")
        self.register_buffer("pos_weight", torch.as_tensor(pos_weight, dtype=torch.float32), persistent=False)
        self.weights = (raw_pred_loss_weight, post_pred_loss_weight, stop_loss_weight)
        self.spectrogram_loss = nn.MSELoss() if spectrogram_loss == "mse" else nn.L1Loss()

    def forward(self, raw_mel_out, postnet_mel_out, stop_out, true_mel, true_stop, true_mask):
        raw = self.spectrogram_loss(raw_mel_out * true_mask, true_mel)
        post = self.spectrogram_loss(postnet_mel_out * true_mask, true_mel)
        stop = binary_cross_entropy_with_logits(stop_out, true_stop, pos_weight=self.pos_weight)
        total = raw * self.weights[0] + post * self.weights[1] + stop * self.weights[2]
        return total, raw, post, stop
