"""Mixed-precision optimizer wrappers with gradient clipping for the GNMT trainer — the roles of
``pipedream-fork/profiler/translation/seq2seq/train/fp_optimizers.py`` (``Fp16Optimizer`` with
fp32 master weights and dynamic loss scaling, ``Fp32Optimizer``) and its ``fused_norm`` ctypes
hook into libTHC (``seq2seq/utils.py:209-219``).  The norm of the flat gradient buffer is one
launch of our ``sumsq`` kernel (``ops.l2norm``); bf16 is the low-precision type on B200 (fp16's
loss scaling is kept for API parity and for fp16 models)."""
from __future__ import annotations

import math
from typing import Iterable, List

import torch
import torch.nn as nn

from .. import ops


def flat_grad_norm(params: Iterable[nn.Parameter]) -> float:
    """L2 norm over all gradients (one fused kernel per contiguous buffer on the GPU)."""
    total = 0.0
    for p in params:
        if p.grad is not None:
            total += float(ops.l2norm(p.grad.reshape(-1))) ** 2
    return math.sqrt(total)


class Fp32Optimizer:
    def __init__(self, model: nn.Module, grad_clip: float = float("inf")):
        self.model, self.grad_clip = model, grad_clip

    def step(self, loss, optimizer, scheduler=None, update: bool = True):
        loss.backward()
        if not update:
            return None
        norm = flat_grad_norm(self.model.parameters())
        if self.grad_clip != float("inf") and norm > self.grad_clip:
            for p in self.model.parameters():
                if p.grad is not None:
                    p.grad.mul_(self.grad_clip / (norm + 1e-6))
        optimizer.step()
        if scheduler is not None:
            scheduler.step()
        self.model.zero_grad(set_to_none=True)
        return norm


class Fp16Optimizer:
    """Low-precision model (fp16/bf16) + ONE flat fp32 master buffer; the wrapped optimizer must
    be built over ``self.fp32_params``.  Dynamic loss scale: halve on overflow (skip the step),
    double after ``upscale_interval`` clean steps."""

    def __init__(self, model: nn.Module, grad_clip: float = float("inf"), loss_scale: float = 8192.0,
                 dls_downscale: float = 2.0, dls_upscale: float = 2.0, dls_upscale_interval: int = 128):
        self.model, self.grad_clip = model, grad_clip
        self.params: List[nn.Parameter] = [p for p in model.parameters() if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.master = torch.empty(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            self.master[off:off + p.numel()].copy_(p.detach().reshape(-1).float())
            off += p.numel()
        self.master = nn.Parameter(self.master)
        self.master.grad = torch.zeros_like(self.master)
        self.fp32_params = [self.master]
        self.loss_scale = loss_scale if self.params[0].dtype == torch.float16 else 1.0
        self.down, self.up, self.interval = dls_downscale, dls_upscale, dls_upscale_interval
        self.since_last_invalid = 0

    def step(self, loss, optimizer, scheduler=None, update: bool = True):
        (loss * self.loss_scale).backward()
        if not update:
            return None
        off = 0
        for p in self.params:
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            self.master.grad[off:off + p.numel()].copy_(g.reshape(-1))
            off += p.numel()
        if self.loss_scale != 1.0:
            self.master.grad.div_(self.loss_scale)
        norm = float(ops.l2norm(self.master.grad))
        if math.isfinite(norm):
            if self.grad_clip != float("inf") and norm > self.grad_clip:
                self.master.grad.mul_(self.grad_clip / (norm + 1e-6))
            optimizer.step()
            if scheduler is not None:
                scheduler.step()
            off = 0
            with torch.no_grad():
                for p in self.params:
                    p.copy_(self.master[off:off + p.numel()].view_as(p))
                    off += p.numel()
            self.since_last_invalid += 1
        else:
            self.loss_scale /= self.down
            self.since_last_invalid = 0
        if self.since_last_invalid >= self.interval and self.loss_scale != 1.0:
            self.loss_scale = min(self.loss_scale * self.up, 8192.0)
            self.since_last_invalid = 0
        self.model.zero_grad(set_to_none=True)
        return norm
