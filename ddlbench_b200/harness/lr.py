"""Learning-rate schedules of the reference scripts.

* imagenet_pytorch: x0.1 every 30 epochs (``benchmark/imagenet/imagenet_pytorch.py:225-229``)
* imagenet_horovod: lr = 0.0125 x world, linear per-batch warm-up during epoch 0 from the
  single-worker LR, then x0.1 at epochs 30/60/80 (``imagenet_horovod.py:262-275``)
* cifar/mnist: constant (x world for horovod)
* pipedream: step / polynomial / exponential with optional 5-epoch warm-up, scaled per stage
  (``pipedream-fork/runtime/image_classification/main_with_runtime.py:605-636``)
"""
from __future__ import annotations

import math

import torch

from ..config import RunConfig


def lr_for(cfg: RunConfig, epoch: int, batch_idx: int, steps_per_epoch: int, world: int = 1) -> float:
    base = cfg.lr
    if cfg.framework == "horovod":
        if cfg.benchmark in ("imagenet", "highres"):
            warm = float(cfg.extra.get("warmup_epochs", 1.0))       # --warmup-epochs (imagenet_horovod.py:52)
            if epoch < warm:
                e = epoch + float(batch_idx + 1) / max(steps_per_epoch, 1)
                adj = 1.0 / world * (e * (world - 1) / warm + 1)
            elif epoch < 30:
                adj = 1.0
            elif epoch < 60:
                adj = 1e-1
            elif epoch < 80:
                adj = 1e-2
            else:
                adj = 1e-3
            return base * world * adj
        return base * world
    if cfg.framework == "pytorch" and cfg.benchmark in ("imagenet", "highres"):
        return base * (0.1 ** (epoch // 30))
    return base


def pipedream_lr(base_lr: float, policy: str, epoch: int, total_epochs: int, step: int,
                 steps_per_epoch: int, warmup_epochs: int = 0, lr_scale: float = 1.0) -> float:
    """PipeDream harness schedule; ``lr_scale`` = replicas_in_stage / replicas_in_first_stage
    (``pipedream-fork/runtime/runtime.py:695-701``)."""
    stage_base = base_lr * lr_scale
    if warmup_epochs > 0 and epoch < warmup_epochs:
        it = step + epoch * steps_per_epoch
        start = 0.1
        frac = float(it) / float(max(warmup_epochs * steps_per_epoch, 1))
        return stage_base * (start + frac * (1.0 - start))
    if policy == "step":
        return stage_base * (0.1 ** (epoch // 30))
    if policy == "polynomial":
        dec = max(1.0 - float(epoch) / float(max(total_epochs, 1)), 0.0) ** 2
        return stage_base * dec
    if policy == "exponential":
        return stage_base * (0.97 ** epoch)
    return stage_base


class WarmupMultiStepLR(torch.optim.lr_scheduler.LRScheduler):
    """GNMT learning-rate policy (``profiler/translation/seq2seq/train/lr_scheduler.py:6-47``):
    exponential warm-up over ``warmup_steps`` iterations (from 1% of the base rate), then the
    base rate until ``remain_steps``, then a decay by ``decay_factor`` every ``decay_interval``
    iterations, at most ``decay_steps`` times."""

    def __init__(self, optimizer, iterations: int, warmup_steps: int = 200, remain_steps=0.666,
                 decay_interval=None, decay_steps: int = 4, decay_factor: float = 0.5, last_epoch=-1):
        def _abs(v):
            return int(v * iterations) if isinstance(v, float) else int(v)
        self.warmup_steps = _abs(warmup_steps)
        self.remain_steps = _abs(remain_steps)
        if decay_interval is None:
            left = max(iterations - self.remain_steps, 0)
            decay_interval = max(left // max(decay_steps, 1), 1)
        self.decay_interval = _abs(decay_interval)
        self.decay_steps, self.decay_factor = decay_steps, decay_factor
        if self.warmup_steps > self.remain_steps:
            self.warmup_steps = self.remain_steps
        super().__init__(optimizer, last_epoch)

    def factor(self, step: int) -> float:
        if step <= self.warmup_steps and self.warmup_steps > 0:
            return math.exp(math.log(0.01) / self.warmup_steps * (self.warmup_steps - step))
        if step >= self.remain_steps:
            n = min((step - self.remain_steps) // self.decay_interval + 1, self.decay_steps)
            return self.decay_factor ** n
        return 1.0

    def get_lr(self):
        f = self.factor(self.last_epoch)
        return [base * f for base in self.base_lrs]
