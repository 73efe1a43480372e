"""Learning-rate schedules of the reference scripts.

* imagenet_pytorch: x0.1 every 30 epochs (``benchmark/imagenet/imagenet_pytorch.py:225-229``)
* imagenet_horovod: lr = 0.0125 x world, linear per-batch warm-up during epoch 0 from the
  single-worker LR, then x0.1 at epochs 30/60/80 (``imagenet_horovod.py:262-275``)
* cifar/mnist: constant (x world for horovod)
* pipedream: step / polynomial / exponential with optional 5-epoch warm-up, scaled per stage
  (``pipedream-fork/runtime/image_classification/main_with_runtime.py:605-636``)
"""
from __future__ import annotations

from ..config import RunConfig


def lr_for(cfg: RunConfig, epoch: int, batch_idx: int, steps_per_epoch: int, world: int = 1) -> float:
    base = cfg.lr
    if cfg.framework == "horovod":
        if cfg.benchmark in ("imagenet", "highres"):
            if epoch < 1:
                e = epoch + float(batch_idx + 1) / max(steps_per_epoch, 1)
                adj = 1.0 / world * (e * (world - 1) / 1.0 + 1)
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
