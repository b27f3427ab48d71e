#!/usr/bin/env python
"""ReformerTTS training-step benchmark (BASELINE.json metric: train mel-frames/s, fwd+bwd, and LSH-attention % of roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config NAME] [--batch B]

Contract (see the task description): one JSON line on rank 0.  A "step" is one full training step of the named reference
config on a synthetic LJSpeech-shaped batch: forward, TTSLoss, backward through the reversible stacks
(with recompute), gradient all-reduce when N > 1, and a fused AdamW update.  ``value`` is measured with the batch resident
in HBM; ``e2e`` repeats the measurement through the public module API with the batch in pinned HOST memory (H2D copies and
the D2H read of the loss inside the timed region).  ``--impl reference`` times the CPU restatement of the reference path
(oracle/model.py - the reference itself cannot travel to the GPU box and reformer_pytorch is not installable) on all host
cores, on a bounded sample of the same workload.  The oracle is only ever the checker / the timed baseline here.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NCU_TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic.json")      # dram__bytes of `ncu --set full` captures, keyed by kernel + shape + source hash
DEFAULT_CONFIG = "bucket-size-64-18-06"       # BASELINE.json configs[1]: the configuration the metric is quoted on
PHONEMES, FRAMES, N_MELS = 200, 800, 80       # "~200 phonemes -> ~800x80 mel frames" (BASELINE.json configs[0])


def synthetic_batch(batch: int, frames: int = FRAMES, phonemes: int = PHONEMES, seed: int = 42, pin: bool = False):
    """Layout of the reference collate function (ref:reformer_tts/dataset/utils.py:5-42), SURVEY.md 8(d)."""
    g = torch.Generator().manual_seed(seed)
    ph = torch.randint(1, 77, (batch, phonemes), generator=g)
    spec = torch.zeros(batch, frames + 1, N_MELS)
    spec[:, 1:] = (torch.randn(batch, frames, N_MELS, generator=g) * 2 - 5).clamp_(-11.5129, 2.0)
    stop = torch.zeros(batch, frames)
    stop[:, -1] = 1
    mask = torch.ones(batch, frames, N_MELS)
    out = {"phonemes": ph, "spectrogram": spec, "stop_tokens": stop, "loss_mask": mask}
    return {k: v.pin_memory() for k, v in out.items()} if pin else out


def train_step(model, loss_fn, optimizer, batch, averager=None):
    """ref:reformer_tts/training/wrappers.py:53-105 (forward + TTSLoss) + backward + AdamW."""
    spec = batch["spectrogram"]
    raw, post, stop, _ = model(batch["phonemes"], spec[:, :-1], batch["loss_mask"].mean(dim=-1))
    loss = loss_fn(raw, post, stop.view(stop.shape[0], -1), spec[:, 1:], batch["stop_tokens"], batch["loss_mask"])[0]
    loss.backward()
    if averager is not None:
        averager.finish()
    optimizer.step()
    optimizer.zero_grad(set_to_none=True)
    return loss


# ------------------------------------------------------------------------------------------------------------------ helpers
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}   # B200_PROFILING.md


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.path = f"/tmp/rtts_clocks_{os.getpid()}.csv"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        self.first = 0

    def mark(self):
        """Samples taken from here on belong to the timed region (the sampler is started before the warm-up: nvidia-smi needs
        ~100 ms to deliver its first line, longer than the timed region of the small configs)."""
        try:
            self.first = sum(1 for _ in open(self.path))
        except OSError:
            self.first = 0

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.proc.wait()
        sm, mx, reasons = [], 0, set()
        lines = open(self.path).readlines()
        in_region = lines[self.first:]
        for line in (in_region if len(in_region) >= 2 else lines):      # (a region shorter than two samples: warm-up samples, same load)
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7 or not f[0].isdigit():
                continue
            sm.append(int(f[0]))
            mx = max(mx, int(f[1]))
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        busy = [c for c in sm if c > 0.5 * mx] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm),
                "samples_in_timed_region": len(in_region)}


def lsh_layer_shapes(kwargs, batch):
    """(name, B, T, R, bucket, D) of the encoder / decoder LSH layers at the padded lengths."""
    pad = kwargs["pad_base"]
    tp, tm = -(-PHONEMES // pad) * pad, -(-FRAMES // pad) * pad
    e, d = kwargs["enc_reformer_kwargs"], kwargs["dec_reformer_kwargs"]
    return {"enc": (batch, tp, e["attn_kwargs"]["n_hashes"], e["attn_kwargs"]["bucket_size"], kwargs["embedding_dim"], e["depth"]),
            "dec": (batch, tm, d["self_attn_kwargs"]["n_hashes"], d["self_attn_kwargs"]["bucket_size"], kwargs["embedding_dim"], d["depth"])}


def workload_config(args, kwargs, world):
    """The workload both arms are measured on (the reference arm times a bounded sample of it: ``cpu_baseline.sample``)."""
    pad = kwargs["pad_base"]
    return {"workload": f"ReformerTTS config/{args.config}.yml training step (fwd + loss + reversible bwd + AdamW)",
            "per_gpu_batch": args.batch, "global_batch": args.batch * world, "phonemes": PHONEMES, "mel_frames": FRAMES,
            "padded_lengths": [-(-PHONEMES // pad) * pad, -(-FRAMES // pad) * pad], "parallelism": f"dp{world}"}


def ncu_traffic(kernel: str, shape: str, sources):
    """DRAM bytes per launch from the committed ncu capture of this kernel at this shape - only if the kernel's source files
    still hash to what was profiled (otherwise the number is stale: None)."""
    from reformer_tts_b200.csrc.build import source_hash
    try:
        entry = json.load(open(NCU_TRAFFIC_FILE)).get(f"{kernel}|{shape}")
    except (OSError, ValueError):
        return None, None
    if not entry or entry.get("source_hash") != source_hash(sources):
        return None, None
    return entry["dram_bytes"], entry.get("profile")


CPU_SAMPLE_FRAMES, CPU_SAMPLE_PHONEMES = 800, 200      # ONE protocol for both CPU legs: one full-length utterance per step


def cpu_sample_description(steps, warmup, cores):
    return (f"1 utterance x {CPU_SAMPLE_FRAMES} mel frames ({CPU_SAMPLE_PHONEMES} phonemes) per step (the workload's batch is "
            f"per_gpu_batch such utterances), {warmup} warm-up + {steps} timed steps, fp32 CPU oracle (oracle/model.py), torch {torch.__version__}, {cores} threads")


# ------------------------------------------------------------------------------------------------------------------ arms
def run_reference(args, kwargs, world, rank):
    """CPU arm: oracle restatement of the reference path, all host threads, bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle.model import ReformerTTSOracle
    from reformer_tts_b200.model.loss import TTSLoss
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(42)
    model = ReformerTTSOracle(**kwargs).train()
    loss_fn = TTSLoss(torch.tensor(5.))
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    # bounded sample of the workload: ONE full-length utterance per step (same protocol as the in-line cpu_baseline leg)
    frames = CPU_SAMPLE_FRAMES
    batch = synthetic_batch(1, frames=frames, phonemes=CPU_SAMPLE_PHONEMES)
    for _ in range(args.warmup):
        train_step(model, loss_fn, opt, batch)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        train_step(model, loss_fn, opt, batch)
    dt = time.perf_counter() - t0
    value = frames * args.steps / dt
    sample = cpu_sample_description(args.steps, args.warmup, cores)
    line = {"impl": "reference", "metric": "train_mel_frames_per_sec", "value": value, "unit": "mel-frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, kwargs, world),
            "cpu_baseline": {"value": value, "unit": "mel-frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "mel-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(kwargs):
    """Rank-0, N=1 only: the oracle port timed beside the GPU number, same bounded sample as ``--impl reference`` (1 warm-up + 2 steps)."""
    from oracle.model import ReformerTTSOracle
    from reformer_tts_b200.model.loss import TTSLoss
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(42)
    model = ReformerTTSOracle(**kwargs).train()
    loss_fn = TTSLoss(torch.tensor(5.))
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    batch = synthetic_batch(1, frames=CPU_SAMPLE_FRAMES, phonemes=CPU_SAMPLE_PHONEMES)
    train_step(model, loss_fn, opt, batch)            # warm-up (first-call overheads)
    t0 = time.perf_counter()
    for _ in range(2):
        train_step(model, loss_fn, opt, batch)
    dt = (time.perf_counter() - t0) / 2
    return {"value": CPU_SAMPLE_FRAMES / dt, "unit": "mel-frames/s", "cores": cores, "kind": "port", "sample": cpu_sample_description(2, 1, cores)}


def run_ours(args, kwargs, world, rank, local_rank):
    import torch.distributed as dist
    from reformer_tts_b200 import _lib, ops
    from reformer_tts_b200.distributed import GradientAverager
    from reformer_tts_b200.model import ReformerTTS
    from reformer_tts_b200.model.loss import TTSLoss
    from reformer_tts_b200.training import TrainStep, make_optimizer, set_lr, warmup_lr
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    _lib.load()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries the ONE result line: whatever NCCL_DEBUG asks NCCL to print while the communicator comes up (its version
        # banner goes to the process's stdout whatever NCCL_DEBUG_FILE says) is sent to stderr instead
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    torch.backends.cuda.matmul.allow_tf32 = True      # non-hot-path torch modules (cross-attention, pre/post nets): fp32 storage, TF32 math
    torch.backends.cudnn.allow_tf32 = True
    torch.manual_seed(42)                               # same initial weights on every rank
    model = ReformerTTS(**kwargs).to(dev).train()
    torch.manual_seed(42 + rank)                        # ref:reformer_tts/training/train.py:16 seeds 42; ranks draw different rotations / dropout
    loss_fn = TTSLoss(torch.tensor(5.)).to(dev)
    # the reference's trainer rules (ref:config/bucket-size-64-18-06.yml:26-32, ref:reformer_tts/training/wrappers.py:240-297): two AdamW
    # parameter groups (no weight decay on biases / LayerNorm gains), lr 3e-4 with a 320-step linear warm-up, global-norm clip 1.0
    base_lr, warmup_steps, grad_clip = 3e-4, 320, 1.0
    opt = make_optimizer(model, base_lr, 1e-6, fused=True, capturable=True)
    averager = GradientAverager(model, overlap=not args.no_overlap) if world > 1 else None
    batch_size = args.batch
    host = synthetic_batch(batch_size, seed=42 + rank, pin=True)
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    # the public training-step API of the package; captures the step in a CUDA graph unless --no-cuda-graph
    step = TrainStep(model, loss_fn, opt, host, use_cuda_graph=not args.no_cuda_graph, averager=averager, seed=1234 + rank,
                     grad_clip=grad_clip, accumulate_grad_batches=args.accumulate)
    resident = step.static                              # device-resident copy of the batch

    def train(batch):
        set_lr(opt, warmup_lr(step.optimizer_steps, base_lr, warmup_steps))      # tensor-valued lr: reaches the replayed graph
        return step.step(batch)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank) if rank == 0 else None      # running from the warm-up on; the timed region is marked below
    for _ in range(args.warmup):
        train(resident)
    sync()
    # ---- timed region 1: inputs resident in HBM ------------------------------------------------------------------------
    shapes = lsh_layer_shapes(kwargs, batch_size)
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    if clocks:
        clocks.mark()
    start.record()
    for _ in range(args.steps):
        train(resident)
    end.record()
    sync()
    ms = start.elapsed_time(end)
    clock_info = clocks.stop() if clocks else None
    # ---- timed region 2: end to end from pinned host memory, loss read back every step -------------------------------------
    sync()
    start.record()
    for _ in range(args.steps):
        loss_host = train(host).item()
    end.record()
    sync()
    ms_e2e = start.elapsed_time(end)
    # ---- per-kernel pass (NOT part of value / e2e): the same step run eagerly with a CUDA-event pair around every launch of
    # libreformer_b200.so on the launching stream; events cannot be recorded inside a graph replay.
    timer = ops.KernelTimer()
    ops.set_kernel_timer(timer)
    launches0 = ops.launch_count()
    start.record()
    for _ in range(2):
        step.eager_step(resident)
    end.record()
    sync()
    eager_ms = start.elapsed_time(end) / 2
    launches_per_step = (ops.launch_count() - launches0) // 2
    kernel_ms = timer.summary()
    ops.set_kernel_timer(None)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    frames_per_step = batch_size * FRAMES * world
    value = frames_per_step * args.steps / (ms / 1e3)
    e2e = frames_per_step * args.steps / (ms_e2e / 1e3)

    if rank == 0:
        peaks = measured_peaks()
        b, t, r, bucket, d, depth = shapes["dec"]
        flops_fwd = 8.0 * b * r * t * bucket * d          # QK^T + PV over a bucket x 2*bucket window, per decoder layer launch
        key_fwd, key_bwd = f"lsh_attn_fwd[T={t}]", f"lsh_attn_bwd[T={t}]"
        fwd_ms, bwd_ms = kernel_ms.get(key_fwd, {}).get("avg_ms"), kernel_ms.get(key_bwd, {}).get("avg_ms")
        peak = peaks["bf16_tflops_sustained"]
        # the forward kernel of the decoder's self-attention: bucket 64 up to T = 2048 runs the paired-chunk kernel (lsh_attn_fwd64p.cu)
        kname, ksrc = (("lsh_attn_fwd64p_kernel", ["lsh_attn_fwd64p.cu", "lsh_attn_params.h", "common.cuh"]) if bucket == 64 and t <= 2048
                       else (f"lsh_attn_fwd_kernel<{bucket}>", ["lsh_attn_fwd.cu", "lsh_attn_params.h", "common.cuh"]))
        roofline = {"bound": "tensor", "kernel": f"{kname} (decoder shape B={b} T={t} H=8 R={r})", "achieved": None, "peak": peak, "unit": "TFLOP/s", "frac": None,
                    "traffic": None, "peak_source": f"{peaks['source']} sustained bf16 (kernel timed inside a long step)"}
        if fwd_ms:
            roofline["achieved"] = flops_fwd / (fwd_ms * 1e-3) / 1e12
            roofline["frac"] = roofline["achieved"] / peak
            roofline["avg_launch_ms"] = fwd_ms
            roofline["launches_timed"] = kernel_ms[key_fwd]["count"]
            roofline["algorithmic_flop_per_launch"] = flops_fwd
            # DRAM bytes of one launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed `ncu --set full` capture of
            # this kernel at this shape; null when the kernel's sources have changed since that capture.  Algorithmic minimum beside it.
            roofline["traffic"], roofline["traffic_source"] = ncu_traffic(kname, f"B={b},T={t},H=8,R={r}", ksrc)
            roofline["algorithmic_min_bytes"] = 2 * b * t * d * 2 + r * b * t * d * 2 + r * b * 8 * t * 4
        if bwd_ms:
            roofline["bwd_kernel"] = {"kernel": f"lsh_attn_bwd_kernel<{bucket}> (decoder shape)", "avg_launch_ms": bwd_ms,
                                      "achieved": 2.5 * flops_fwd / (bwd_ms * 1e-3) / 1e12, "frac": 2.5 * flops_fwd / (bwd_ms * 1e-3) / 1e12 / peak}
        roofline["timed_in"] = "eager pass of the same step, CUDA-event pair around each launch (events cannot be recorded inside a CUDA-graph replay)"
        own_total = sum(v["total_ms"] for v in kernel_ms.values()) / 2
        roofline["share_of_step"] = {k: round(v["total_ms"] / 2 / (ms / args.steps), 4) for k, v in sorted(kernel_ms.items(), key=lambda kv: -kv[1]["total_ms"])[:12]}
        roofline["own_kernels_ms_per_step"] = own_total
        cpu = cpu_baseline_leg(kwargs) if world == 1 and not args.no_cpu_baseline else None
        line = {"metric": "train_mel_frames_per_sec", "value": value, "unit": "mel-frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(args, kwargs, world),
                "run": {"precision": "bf16 MMA operands, fp32 accumulate, fp32 residual stream / master weights; non-hot-path torch modules fp32 storage + TF32",
                        "l2": "no flush: one step streams several GB of activations (>> 126 MB L2) through HBM",
                        "trainer": f"torch AdamW(fused, capturable, tensor lr) over the reference's two parameter groups, lr {base_lr} with {warmup_steps}-step "
                                   f"linear warm-up set every step, global-norm clip {grad_clip} after the all-reduce, accumulate_grad_batches "
                                   f"{args.accumulate} - all inside the timed region",
                        "gradient_allreduce": None if world == 1 else ("per reversible block, overlapped with the reversible backward, inside the captured graph"
                                                                      if averager.overlap else "one flat all-reduce after backward"),
                        "cuda_graph": step.graph is not None, "cuda_graph_error": step.graph_error, "eager_ms_per_step": eager_ms,
                        "build_id": _lib.build_id()},
                "e2e": {"value": e2e, "unit": "mel-frames/s", "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "last_loss": loss_host},
                "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step, "clocks": clock_info, "roofline": roofline}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        # Leave without tearing NCCL down: destroying a communicator whose collectives were captured into a CUDA graph hung the
        # workers after the result line had been printed (2 GPUs, torch 2.11 / NCCL 2.28).  Everyone syncs, flushes and exits.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=DEFAULT_CONFIG)
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: the reference config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cuda-graph", action="store_true", help="run the step eagerly instead of replaying a captured CUDA graph")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: one flat gradient all-reduce after backward instead of per-block overlapped ones")
    ap.add_argument("--accumulate", type=int, default=1, help="accumulate_grad_batches (the reference YAMLs use 3-5; 1 = every step reduces and updates)")
    args = ap.parse_args()
    from reformer_tts_b200.model import config as C
    kwargs = C.reference_model_kwargs(args.config)
    if args.batch is None:
        args.batch = C.REFERENCE_BATCH[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch multi-GPU runs with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N bench.py --gpus N ...")
    if args.impl == "reference":
        run_reference(args, kwargs, world, rank)
    else:
        run_ours(args, kwargs, world, rank, local_rank)


if __name__ == "__main__":
    main()
