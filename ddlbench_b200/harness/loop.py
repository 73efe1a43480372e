"""Epoch loops of the `pytorch` (single GPU) and `horovod` (data parallel) benchmarks with the
reference's log lines and metric definitions (``benchmark/imagenet/imagenet_pytorch.py:123-222``,
``imagenet_horovod.py:163-255``), plus the device-timed measurement BASELINE.json asks for.

Differences from the reference that are deliberate: throughput is timed on
the device with CUDA events and reduced with MAX over ranks instead of per-rank host wall
clock; the loss is read back asynchronously (lagged one step) instead of forcing a sync.
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from ..config import RunConfig
from ..data.pipeline import DevicePrefetcher
from ..data.synthetic import SyntheticImages
from . import logfmt
from .lr import lr_for
from .trainer import Trainer


class StepTimer:
    """CUDA-event (or host, on CPU) timer around a region of the launching stream."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.gpu = self.device.type == "cuda"
        if self.gpu:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
        self.t0 = 0.0

    def start(self):
        if self.gpu:
            torch.cuda.synchronize(self.device)
            self.e0.record(torch.cuda.current_stream(self.device))
        self.t0 = time.perf_counter()

    def stop(self) -> float:
        """Milliseconds."""
        if self.gpu:
            self.e1.record(torch.cuda.current_stream(self.device))
            torch.cuda.synchronize(self.device)
            return self.e0.elapsed_time(self.e1)
        return (time.perf_counter() - self.t0) * 1e3


def max_over_ranks(value: float, device, group=None) -> float:
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        t = torch.tensor([value], dtype=torch.float64,
                         device=device if dist.get_backend(group) == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return float(t.item())
    return value


def timed_loop(trainer: Trainer, pre: DevicePrefetcher, steps: int, warmup: int,
               read_loss: bool = True) -> Tuple[float, List[float]]:
    """End-to-end loop through the public step API: every step copies its uint8 batch from
    pinned host memory (copy stream, overlapped), trains, and reads the loss back to the
    host (asynchronously, collected one step later).  Returns (ms/step device-timed, losses)."""
    dev = trainer.device
    gpu = dev.type == "cuda"
    src = pre.src
    pinned = [torch.zeros(3, dtype=torch.float32).pin_memory() if gpu else torch.zeros(3)
              for _ in range(2)]
    evs: List[Optional[torch.cuda.Event]] = [None, None]
    losses: List[float] = []
    timer = StepTimer(dev)
    pre.issue()
    total = warmup + steps
    for i in range(total):
        if i == warmup:
            if gpu:
                torch.cuda.synchronize(dev)
            if trainer.distributed:
                dist.barrier(trainer.group)
            timer.start()
        if i + 1 < total:
            pre.issue()
        x_u8, tgt = pre.get(i)
        stats = trainer.step_u8(x_u8, tgt, src.mean, src.std)
        pre.release(i)
        if read_loss:
            k = i & 1
            if evs[k] is not None:
                evs[k].synchronize()
                losses.append(float(pinned[k][0]))
            pinned[k].copy_(stats, non_blocking=True)
            if gpu:
                evs[k] = torch.cuda.Event()
                evs[k].record(torch.cuda.current_stream(dev))
    if read_loss and gpu:
        for k in ((total) & 1, (total + 1) & 1):
            if evs[k] is not None:
                evs[k].synchronize()
                losses.append(float(pinned[k][0]))
    ms = timer.stop()
    if trainer.distributed:
        dist.barrier(trainer.group)
    return ms / max(steps, 1), losses


def run(cfg: RunConfig, group=None, verbose: bool = True) -> Dict[str, float]:
    """Run the `pytorch` or `horovod` benchmark described by ``cfg``; returns the final
    metrics.  Prints the reference's log lines on every rank that the reference prints on."""
    distributed = cfg.framework == "horovod" and dist.is_initialized() and dist.get_world_size(group) > 1
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    use_cuda = cfg.device != "cpu" and torch.cuda.is_available()
    device = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    say = print if verbose else (lambda *a, **k: None)
    say(logfmt.creating_model(cfg.arch))
    base_lr = cfg.lr * (world if cfg.framework == "horovod" else 1)
    trainer = Trainer(cfg.arch, cfg.benchmark, cfg.batch_size, device=str(device),
                      dtype=(torch.bfloat16 if (use_cuda and cfg.dtype == "bf16") else torch.float32),
                      lr=base_lr, momentum=cfg.momentum, weight_decay=cfg.weight_decay,
                      cuda_graph=cfg.cuda_graph and use_cuda, distributed=distributed,
                      allreduce=cfg.allreduce, tcgen05=cfg.fused and use_cuda, seed=cfg.seed,
                      group=group)
    pinned = cfg.data_mode == "pinned" and use_cuda
    from ..data import make_source
    train_src = make_source(cfg, cfg.batch_size, device, rank, world, "train", pinned=pinned)
    real = not isinstance(train_src, SyntheticImages)
    val_src = make_source(cfg, cfg.batch_size, device, rank, world, "val") if real else train_src
    # the image-tree source is already sharded per rank; the synthetic one models the full set
    steps = len(train_src) if real else (len(train_src) // world if distributed else len(train_src))
    if cfg.steps_per_epoch:
        steps = min(steps, cfg.steps_per_epoch)
    val_steps = cfg.eval_steps if cfg.eval_steps is not None else max(1, steps // 10)
    pre = DevicePrefetcher(train_src, device)
    throughputs, elapsed, dev_tputs = [], [], []
    valid_acc = 0.0
    for epoch in range(cfg.epochs):
        timer = StepTimer(device)
        tick = time.time()
        timer.start()
        loss_sum = torch.zeros(1, device=device)
        data_trained = 0
        if real:
            train_src.set_epoch(epoch)
        pre.i_issue = 0
        pre.issue()
        for i in range(steps):
            trainer.set_lr(lr_for(cfg, epoch, i, steps, world))
            if i + 1 < steps:
                pre.issue()
            x_u8, tgt = pre.get(i)
            stats = trainer.step_u8(x_u8, tgt, train_src.mean, train_src.std)
            pre.release(i)
            data_trained += cfg.batch_size
            loss_sum += stats[0].detach() * cfg.batch_size
            if i % cfg.log_interval == 0:
                say(logfmt.train_progress(epoch + 1, cfg.epochs, i / steps * 100,
                                          data_trained / max(time.time() - tick, 1e-9), device))
        ms = timer.stop()
        tock = time.time()
        train_loss = float(loss_sum.item()) / max(data_trained, 1)
        vloss, vacc = evaluate(trainer, val_src, min(val_steps, len(val_src)), cfg, say)
        t = tock - tick
        thr = data_trained / max(t, 1e-9)
        ms_max = max_over_ranks(ms, device, group)
        dev_tputs.append(data_trained * world / (ms_max / 1e3))
        say(logfmt.epoch_summary(epoch + 1, cfg.epochs, train_loss, thr, vloss, vacc))
        say(logfmt.device_timed(dev_tputs[-1], ms_max / max(steps, 1), world))
        throughputs.append(thr)
        elapsed.append(t)
        valid_acc = vacc
    vloss, valid_acc = evaluate(trainer, val_src, min(val_steps, len(val_src)), cfg, lambda *a, **k: None)
    if cfg.framework == "horovod":
        from ..parallel import dp
        vloss = float(dp.allreduce(torch.tensor(vloss), name="avg_loss"))
        valid_acc = float(dp.allreduce(torch.tensor(valid_acc), name="avg_accuracy"))
        if rank == 0:
            say(logfmt.test_set(vloss, 100.0 * valid_acc))
    n = max(len(throughputs), 1)
    out = {"throughput": sum(throughputs) / n, "sec_per_epoch": sum(elapsed) / n,
           "device_throughput": sum(dev_tputs) / n, "valid_accuracy": valid_acc,
           "peak_mem_gb": logfmt.memory_gb(device if use_cuda else None)[0]}
    say(logfmt.final_summary(valid_acc, out["throughput"], out["sec_per_epoch"]))
    trainer.check()
    return out


def evaluate(trainer: Trainer, src: SyntheticImages, steps: int, cfg: RunConfig, say):
    tick = time.time()
    loss_sum = torch.zeros(1, device=trainer.device)
    correct = torch.zeros(1, device=trainer.device)
    n = 0
    from .. import ops
    for i in range(steps):
        img, lab = src.batch(i)
        img = img.to(trainer.device, non_blocking=True)
        lab = lab.to(trainer.device, non_blocking=True)
        x = ops.normalize_u8(img, src.mean, src.std, dtype=trainer.dtype)
        stats = trainer.eval_step(x, lab)
        loss_sum += stats[0] * lab.shape[0]
        correct += stats[1]
        n += lab.shape[0]
        if i % cfg.log_interval == 0:
            say(logfmt.valid_progress(i / steps * 100, n / max(time.time() - tick, 1e-9)))
    return float(loss_sum.item()) / max(n, 1), float(correct.item()) / max(n, 1)
