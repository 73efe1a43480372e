"""Dump per-layer activations and gradients every k epochs for offline sparsity /
compression studies — the role of ``torchlogger.ActivationAndGradientLogger``
(``pipedream-fork/profiler/torchmodules/torchlogger/activation_gradient_logger.py:9-78``).

Works on any ``nn.Module`` through forward hooks and full-backward hooks (stock PyTorch has
them; the reference needed its autograd patch for backward PRE-hooks only)."""
from __future__ import annotations

import os
import pickle
from typing import Dict, List

import torch
import torch.nn as nn


class ActivationAndGradientLogger:
    def __init__(self, directory: str, max_elements: int = 1 << 22):
        self.directory = directory
        self.max_elements = max_elements
        os.makedirs(directory, exist_ok=True)
        self.activations: Dict[str, torch.Tensor] = {}
        self.gradients: Dict[str, torch.Tensor] = {}
        self._handles: List = []

    def _keep(self, t: torch.Tensor) -> torch.Tensor:
        t = t.detach().float().reshape(-1)
        return t[: self.max_elements].cpu()

    def hook_modules(self, module: nn.Module, prefix: str = "") -> None:
        for name, sub in module.named_modules():
            if list(sub.children()):
                continue
            key = (prefix + name) or type(sub).__name__

            def fwd(m, inp, out, key=key):
                if torch.is_tensor(out):
                    self.activations[key] = self._keep(out)

            def bwd(m, gin, gout, key=key):
                if gout and torch.is_tensor(gout[0]):
                    self.gradients[key] = self._keep(gout[0])

            self._handles.append(sub.register_forward_hook(fwd))
            self._handles.append(sub.register_full_backward_hook(bwd))

    def unhook_modules(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []

    def reset(self) -> None:
        self.activations, self.gradients = {}, {}

    def dump(self, epoch: int) -> None:
        d = os.path.join(self.directory, str(epoch))
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "activations.pkl"), "wb") as f:
            pickle.dump(self.activations, f)
        with open(os.path.join(d, "gradients.pkl"), "wb") as f:
            pickle.dump(self.gradients, f)
