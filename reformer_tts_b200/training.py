"""One ReformerTTS training step as a callable: forward, TTSLoss, reversible backward, gradient averaging, global-norm
clipping, optimiser update (ref:reformer_tts/training/wrappers.py:53-105,234-297 + the Lightning loop configured at
ref:reformer_tts/training/train.py:77-89: ``accumulate_grad_batches``, ``gradient_clip_val``).  The whole step can be
captured ONCE in a CUDA graph and replayed: at the reference configs a step is ~1250 kernel launches of 2-420 us, so eager
launch overhead is as large as the GPU time itself.  Capture needs static shapes (the collate pads to a fixed length),
capture-safe randomness (``Deterministic.use_private_generators``) and an optimiser whose hyper-parameters live on the
device (``make_optimizer(..., capturable=True)``: tensor-valued ``lr``, so the reference's warm-up reaches the replayed graph).

Gradients live in one flat buffer (``distributed.GradientBuckets``): the data-parallel all-reduce runs per reversible block
during the backward pass, inside the captured graph too; clipping and zeroing are single kernels over that buffer."""
from __future__ import annotations

import sys
from typing import Dict, Optional

import torch
from torch import nn

from .distributed import GradientBuckets
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


def make_optimizer(model: nn.Module, learning_rate: float, weight_decay: float, fused: Optional[bool] = None,
                   capturable: Optional[bool] = None) -> torch.optim.AdamW:
    """AdamW over ``param_groups`` (ref:...wrappers.py:251-256).  ``fused`` and ``capturable`` default to True on CUDA parameters.
    A capturable optimiser keeps ``lr`` as a device tensor: ``set_lr`` then changes the rate a captured CUDA graph applies."""
    on_cuda = next(model.parameters()).is_cuda
    if fused is None:
        fused = on_cuda
    if capturable is None:
        capturable = on_cuda
    lr = torch.tensor(float(learning_rate), dtype=torch.float32, device=next(model.parameters()).device) if capturable else learning_rate
    return torch.optim.AdamW(param_groups(model, weight_decay), lr=lr, weight_decay=weight_decay, fused=fused, capturable=capturable)


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


def clip_coefficient(total_norm: torch.Tensor, max_norm: float) -> torch.Tensor:
    """``torch.nn.utils.clip_grad_norm_``'s factor (what Lightning's ``gradient_clip_val`` applies, ref:...train.py:88):
    max_norm / (total_norm + 1e-6), clamped to 1."""
    return (max_norm / (total_norm + 1e-6)).clamp(max=1.0)


def loss_of_batch(model, loss_fn, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
    spec = batch["spectrogram"]
    raw, post, stop, _ = model(batch["phonemes"], spec[:, :-1], batch["loss_mask"].mean(dim=-1))
    return loss_fn(raw, post, stop.view(stop.shape[0], -1), spec[:, 1:], batch["stop_tokens"], batch["loss_mask"])[0]


class TrainStep:
    """``step(batch) -> loss`` (a 0-dim device tensor).  ``batch`` tensors may live on the host (pinned) or on the device;
    with ``use_cuda_graph`` they are copied into static device buffers and a captured graph is replayed.

    ``accumulate_grad_batches = k``: every call is one micro-batch (loss / k back-propagated into the flat gradient buffer);
    the k-th call also averages over the ranks, clips to ``grad_clip`` (global 2-norm, after the all-reduce), applies the
    optimiser and zeroes the buffer.  Micro-batches and boundary steps are two captured graphs.
    ``optimizer_steps`` counts boundary steps; the caller sets the rate with ``set_lr(optimizer, warmup_lr(...))``."""

    def __init__(self, model: nn.Module, loss_fn: nn.Module, optimizer: torch.optim.Optimizer, example_batch: Dict[str, torch.Tensor],
                 use_cuda_graph: bool = True, averager=None, warmup_steps: int = 3, seed: int = 1234, grad_clip: Optional[float] = None,
                 accumulate_grad_batches: int = 1, private_rng: Optional[bool] = None, keep_grads: bool = False):
        self.model, self.loss_fn, self.optimizer, self.averager = model, loss_fn, optimizer, averager
        self.grad_clip, self.accumulate = grad_clip, int(accumulate_grad_batches)
        assert self.accumulate >= 1
        self.device = next(model.parameters()).device
        self.buckets: GradientBuckets = averager.buckets if averager is not None else GradientBuckets(model)
        self.static = {k: v.to(self.device).clone() for k, v in example_batch.items()}
        self.graph: Optional[torch.cuda.CUDAGraph] = None             # boundary step (reduce + clip + update)
        self.graph_micro: Optional[torch.cuda.CUDAGraph] = None       # non-boundary micro-batch (accumulate only)
        self.static_loss = self.static_loss_micro = None
        self.graph_error = None
        self.grad_norm: Optional[torch.Tensor] = None                  # global gradient norm of the latest boundary step (before clipping)
        self.keep_grads = keep_grads                                   # tests: keep a copy of the (averaged, clipped) flat gradient of the latest boundary step
        self.last_grads: Optional[torch.Tensor] = None
        self.optimizer_steps = 0
        self._micro_index = 0
        self._generators = []
        # capture-safe randomness of the reversible recompute (mandatory under capture, optional in eager mode - the default there is
        # the reference's record / replay of the global generator state)
        if use_cuda_graph:
            self._check_capturable()
        if use_cuda_graph or private_rng:
            for i, mod in enumerate(m for m in self.model.modules() if isinstance(m, Deterministic)):
                self._generators.append(mod.use_private_generators(seed + i, self.device))
        if use_cuda_graph:
            rng_before = torch.cuda.get_rng_state(self.device)
            try:
                self._capture(warmup_steps, seed)
            except Exception as exc:      # capture is an optimisation: report and run eagerly
                self.graph, self.graph_micro, self.graph_error = None, None, f"{type(exc).__name__}: {exc}"
                print(f"[reformer_tts_b200] CUDA-graph capture failed, running eagerly: {self.graph_error}", file=sys.stderr)
                torch.cuda.synchronize()
                # an aborted capture leaves the generators in capture mode: go back to the default record / replay RNG handling
                for mod in self.model.modules():
                    if isinstance(mod, Deterministic):
                        mod._private = None
                self._generators = []
                torch.cuda.set_rng_state(rng_before, self.device)
                self.buckets.zero()

    # ------------------------------------------------------------------------------------------------------------------
    def _check_capturable(self):
        """A captured ``optimizer.step()`` needs ``capturable=True`` (torch refuses otherwise) and a tensor-valued lr (a float
        would be baked into the graph: the reference's 320-step warm-up would silently do nothing on replay)."""
        for group in self.optimizer.param_groups:
            if not group.get("capturable", False):
                raise ValueError("TrainStep(use_cuda_graph=True) needs an optimiser built with capturable=True (training.make_optimizer does)")
            if not isinstance(group["lr"], torch.Tensor):
                raise ValueError("TrainStep(use_cuda_graph=True) needs a tensor-valued lr (training.make_optimizer(capturable=True)) so that "
                                 "set_lr reaches the captured graph")

    def _run(self, batch, boundary: bool):
        """One micro-batch; ``boundary``: also reduce, clip, update, zero."""
        if not self.buckets.attached():
            raise RuntimeError("parameter gradients were detached from the flat buffer (do not call zero_grad(set_to_none=True))")
        if self._weights_changed:
            _WeightCache.epoch += 1       # bf16 weight copies are rebuilt once per optimiser step (inside the captured graph too),
            _WeightCache.refresh_all()    # all of them by one multi-tensor cast
        if self.averager is not None:
            self.averager.sync_enabled = boundary
        loss = loss_of_batch(self.model, self.loss_fn, batch)
        (loss if self.accumulate == 1 else loss * (1.0 / self.accumulate)).backward()
        if boundary:
            if self.averager is not None:
                self.averager.finish()
            if self.grad_clip is not None:      # global 2-norm of the AVERAGED gradient: after the all-reduce
                self.grad_norm = torch.linalg.vector_norm(self.buckets.flat)
                self.buckets.flat.mul_(clip_coefficient(self.grad_norm, self.grad_clip))
            if self.keep_grads:
                if self.last_grads is None:
                    self.last_grads = torch.empty_like(self.buckets.flat)
                self.last_grads.copy_(self.buckets.flat)
            self.optimizer.step()
            self.buckets.zero()
        return loss

    def _snapshot(self):
        model_state = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        opt_state = {id(p): {k: v.detach().clone() for k, v in st.items() if isinstance(v, torch.Tensor)} for p, st in self.optimizer.state.items()}
        lrs = [g["lr"].clone() if isinstance(g["lr"], torch.Tensor) else g["lr"] for g in self.optimizer.param_groups]
        return model_state, opt_state, lrs

    def _restore(self, snap):
        """In place: the captured graph refers to the storage the warm-up allocated (lazily created optimiser state included)."""
        model_state, opt_state, lrs = snap
        with torch.no_grad():
            for k, v in self.model.state_dict().items():
                v.copy_(model_state[k])
            for p, st in self.optimizer.state.items():
                before = opt_state.get(id(p), {})
                for k, v in st.items():
                    if isinstance(v, torch.Tensor):
                        if k in before:
                            v.copy_(before[k])
                        else:
                            v.zero_()       # state that did not exist before the warm-up: AdamW creates zeros / step 0
            for g, lr in zip(self.optimizer.param_groups, lrs):
                if isinstance(g["lr"], torch.Tensor):
                    g["lr"].copy_(lr)
                else:
                    g["lr"] = lr
        self.buckets.zero()

    def reseed(self, seed: int) -> None:
        """Restart the private generator pairs (sub-network i: seed + i).  Two TrainSteps over equal models, reseeded alike, draw
        the same rotations and dropout masks whether they replay a graph or run eagerly."""
        for i, (fwd, rec) in enumerate(self._generators):
            fwd.manual_seed(seed + i)
            rec.manual_seed(seed + i)

    def _capture(self, warmup_steps: int, seed: int):
        generators = [g for pair in self._generators for g in pair]
        # Warm-up on a side stream (allocator, cuDNN / NCCL initialisation, lazily created optimiser state).  It runs real
        # optimiser updates on the example batch, so weights, buffers, optimiser state and learning rate are put back afterwards:
        # constructing a TrainStep does not train.
        snap = self._snapshot()
        self._weights_changed = True
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup_steps):
                for m in range(self.accumulate):
                    self._run(self.static, boundary=m == self.accumulate - 1)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._restore(snap)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        for g in generators:
            graph.register_generator_state(g)
        # the epoch bump inside _run makes the first use of every weight in the captured step re-cast it, so the fp32 -> bf16
        # casts are recorded in the graph once per weight and replayed after every optimiser update
        with torch.cuda.graph(graph):
            self.static_loss = self._run(self.static, boundary=True)
        self.graph = graph
        if self.accumulate > 1:
            micro = torch.cuda.CUDAGraph()
            for g in generators:
                micro.register_generator_state(g)
            self._weights_changed = True      # the first micro-batch after an update re-casts; later ones repeat the (idempotent) cast
            with torch.cuda.graph(micro, pool=graph.pool()):
                self.static_loss_micro = self._run(self.static, boundary=False)
            self.graph_micro = micro
        self.buckets.zero()

    # ------------------------------------------------------------------------------------------------------------------
    _weights_changed = True

    def _advance(self) -> bool:
        """-> is this call the boundary micro-batch?"""
        self._micro_index += 1
        if self._micro_index == self.accumulate:
            self._micro_index = 0
            self.optimizer_steps += 1
            return True
        return False

    def eager_step(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        """The same step without the graph (profiling passes that need per-kernel events)."""
        on_dev = {k: (v if v.device == self.device else v.to(self.device, non_blocking=True)) for k, v in batch.items()}
        boundary = self._advance()
        loss = self._run(on_dev, boundary)
        self._weights_changed = boundary
        return loss

    def step(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        if self.graph is None:
            return self.eager_step(batch)
        for k, v in batch.items():
            if v.data_ptr() != self.static[k].data_ptr():
                self.static[k].copy_(v, non_blocking=True)
        if self._advance():
            self.graph.replay()
            return self.static_loss
        self.graph_micro.replay()
        return self.static_loss_micro
