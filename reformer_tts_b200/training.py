"""One ReformerTTS training step as a callable: forward, TTSLoss, reversible backward, gradient averaging, optimiser update
(ref:reformer_tts/training/wrappers.py:53-105 + the Lightning loop around it).  The whole step can be captured ONCE in a CUDA
graph and replayed: at the reference configs a step is ~1250 kernel launches of 2-420 us, so eager launch overhead is as large
as the GPU time itself.  Capture needs static shapes (the collate pads to a fixed length) and capture-safe randomness
(``Deterministic.use_private_generators``)."""
from __future__ import annotations

import sys
from typing import Dict, Optional

import torch
from torch import nn

from .lsh_attention import _WeightCache
from .model.reversible import Deterministic


NO_DECAY = ("bias", "norm.weight")      # ref:reformer_tts/training/wrappers.py:240 ("norm.weight only applies to nn.LayerNorm")


def param_groups(model: nn.Module, weight_decay: float):
    """The reference's two AdamW parameter groups (ref:reformer_tts/training/wrappers.py:240-250): every parameter whose NAME contains
    "bias" or "norm.weight" is exempt from weight decay.  The rule keys on names, which is why the product modules keep the
    reference's parameter names."""
    decay = [p for n, p in model.named_parameters() if not any(nd in n for nd in NO_DECAY)]
    no_decay = [p for n, p in model.named_parameters() if any(nd in n for nd in NO_DECAY)]
    return [{"params": decay, "weight_decay": weight_decay}, {"params": no_decay, "weight_decay": 0.0}]


def make_optimizer(model: nn.Module, learning_rate: float, weight_decay: float, fused: Optional[bool] = None) -> torch.optim.AdamW:
    """AdamW over ``param_groups`` (ref:...wrappers.py:251-256).  ``fused`` defaults to True on CUDA parameters."""
    if fused is None:
        fused = next(model.parameters()).is_cuda
    return torch.optim.AdamW(param_groups(model, weight_decay), lr=learning_rate, weight_decay=weight_decay, fused=fused)


def warmup_lr(global_step: int, base_lr: float, warmup_steps: Optional[int]) -> float:
    """Learning rate of optimiser step ``global_step`` (0-based) under the reference's linear warm-up
    (ref:...wrappers.py:286-295): base_lr * min(1, (step + 1) / warmup_steps) while step < warmup_steps."""
    if warmup_steps is None or global_step >= warmup_steps:
        return base_lr
    return min(1.0, float(global_step + 1) / warmup_steps) * base_lr


def set_lr(optimizer: torch.optim.Optimizer, lr: float) -> None:
    """Write ``lr`` into every parameter group; a tensor-valued ``lr`` (CUDA-graph captured optimisers) is updated in place."""
    for group in optimizer.param_groups:
        if isinstance(group["lr"], torch.Tensor):
            group["lr"].fill_(lr)
        else:
            group["lr"] = lr


def loss_of_batch(model, loss_fn, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
    spec = batch["spectrogram"]
    raw, post, stop, _ = model(batch["phonemes"], spec[:, :-1], batch["loss_mask"].mean(dim=-1))
    return loss_fn(raw, post, stop.view(stop.shape[0], -1), spec[:, 1:], batch["stop_tokens"], batch["loss_mask"])[0]


class TrainStep:
    """``step(batch) -> loss`` (a 0-dim device tensor).  ``batch`` tensors may live on the host (pinned) or on the device;
    with ``use_cuda_graph`` they are copied into static device buffers and the captured graph is replayed."""

    def __init__(self, model: nn.Module, loss_fn: nn.Module, optimizer: torch.optim.Optimizer, example_batch: Dict[str, torch.Tensor],
                 use_cuda_graph: bool = True, averager=None, warmup_steps: int = 3, seed: int = 1234):
        self.model, self.loss_fn, self.optimizer, self.averager = model, loss_fn, optimizer, averager
        self.device = next(model.parameters()).device
        self.static = {k: v.to(self.device).clone() for k, v in example_batch.items()}
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.static_loss = None
        self.graph_error = None
        if use_cuda_graph:
            rng_before = torch.cuda.get_rng_state(self.device)
            try:
                self._capture(warmup_steps, seed)
            except Exception as exc:      # capture is an optimisation: report and run eagerly
                self.graph, self.graph_error = None, f"{type(exc).__name__}: {exc}"
                print(f"[reformer_tts_b200] CUDA-graph capture failed, running eagerly: {self.graph_error}", file=sys.stderr)
                torch.cuda.synchronize()
                # an aborted capture leaves the generators in capture mode: go back to the default record / replay RNG handling
                for mod in self.model.modules():
                    if isinstance(mod, Deterministic):
                        mod._private = None
                torch.cuda.set_rng_state(rng_before, self.device)
                if self.averager is not None and hasattr(self.averager, "enable_overlap"):
                    self.averager.enable_overlap()

    # ------------------------------------------------------------------------------------------------------------------
    def _eager(self, batch):
        _WeightCache.epoch += 1       # bf16 weight copies are rebuilt once per step (inside the captured graph too)
        loss = loss_of_batch(self.model, self.loss_fn, batch)
        loss.backward()
        if self.averager is not None:
            self.averager.finish()
        self.optimizer.step()
        return loss

    def _capture(self, warmup_steps: int, seed: int):
        if self.averager is not None:
            self.averager.disable_overlap()
        generators = []
        for i, mod in enumerate(m for m in self.model.modules() if isinstance(m, Deterministic)):
            generators += mod.use_private_generators(seed + i, self.device)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup_steps):
                self.optimizer.zero_grad(set_to_none=True)
                self._eager(self.static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.optimizer.zero_grad(set_to_none=True)
        graph = torch.cuda.CUDAGraph()
        for g in generators:
            graph.register_generator_state(g)
        # the epoch bump inside _eager makes the first use of every weight in the captured step re-cast it, so the fp32 -> bf16
        # casts are recorded in the graph once per weight and replayed after every optimiser update
        with torch.cuda.graph(graph):
            self.static_loss = self._eager(self.static)
        self.graph = graph

    # ------------------------------------------------------------------------------------------------------------------
    def eager_step(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        """The same step without the graph (profiling passes that need per-kernel events)."""
        on_dev = {k: (v if v.device == self.device else v.to(self.device, non_blocking=True)) for k, v in batch.items()}
        self.optimizer.zero_grad(set_to_none=True)
        return self._eager(on_dev)

    def step(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        if self.graph is None:
            on_dev = {k: (v if v.device == self.device else v.to(self.device, non_blocking=True)) for k, v in batch.items()}
            self.optimizer.zero_grad(set_to_none=True)
            return self._eager(on_dev)
        for k, v in batch.items():
            if v.data_ptr() != self.static[k].data_ptr():
                self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.static_loss
