"""Weight-stashing optimizers for plain ``nn.Module`` stages (the autograd path).

API-compatible with the reference's ``OptimizerWithWeightStashing`` /
``SGDWithWeightStashing`` / ``AdamWithWeightStashing``
(``pipedream-fork/runtime/{optimizer,sgd,adam}.py``): ``load_old_params()`` before backward,
``load_new_params()`` before ``step()``, a ring of ``num_versions`` weight versions,
``macrobatch`` mode (<= 2 versions, one step per ``num_versions`` minibatches with averaged
gradients), optional static loss scaling with fp32 master parameters.

Mechanism: every parameter of the wrapped modules is re-homed as a VIEW into one flat buffer
per version; switching version re-points ``param.data`` (no ``state_dict`` clones, no
``load_state_dict`` — the reference performs two full loads and one full copy per minibatch,
``optimizer.py:110-116,157-159``).  BatchNorm running statistics are ordinary buffers and are
never versioned (reference ``optimizer.py:76-79``).  The engine-based ``StageRuntime`` does not
need this class (its ``ParamStore`` already has version slots); it exists for generated /
user-written ``nn.Module`` stages and for API parity.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.nn as nn


class Version:
    def __init__(self, version: int = 0):
        self.version = version

    def incr(self) -> "Version":
        return Version(self.version + 1)

    def __repr__(self):
        return "v%d" % self.version


class OptimizerWithWeightStashing:
    def __init__(self, optim_name: str, modules: Sequence[nn.Module], master_parameters=None,
                 model_parameters=None, loss_scale: float = 1.0, num_versions: int = 1,
                 verbose_freq: int = 0, macrobatch: bool = False, **optimizer_args):
        self.modules = list(modules)
        self.loss_scale = float(loss_scale)
        self.verbose_freq = verbose_freq
        self.batch_counter = 0
        self.macrobatch = macrobatch
        self.update_interval = 1
        if macrobatch:
            self.update_interval = max(num_versions, 1)
            num_versions = min(2, num_versions)
        self.num_versions = max(1, num_versions)
        # ---- re-home parameters into flat version buffers
        self.params: List[nn.Parameter] = []
        seen = set()
        for m in self.modules:
            for p in m.parameters():
                if id(p) not in seen:
                    seen.add(id(p))
                    self.params.append(p)
        total = sum(p.numel() for p in self.params)
        ref = self.params[0] if self.params else torch.zeros(1)
        self.flat = [torch.empty(total, dtype=ref.dtype, device=ref.device)
                     for _ in range(self.num_versions)]
        self.views: List[List[torch.Tensor]] = []
        for v in range(self.num_versions):
            off, vs = 0, []
            for p in self.params:
                n = p.numel()
                view = self.flat[v][off:off + n].view_as(p)
                view.copy_(p.data)
                vs.append(view)
                off += n
            self.views.append(vs)
        self.latest_version = Version()
        self.current_version = Version()
        self._slot_of_version: Dict[int, int] = {0: 0}
        self._latest_slot = 0
        self._point_to(0)
        # The wrapped optimizer updates SHADOW parameters that alias the target slot: in-place
        # updates then never bump the autograd version counter of the model's own parameters
        # (modern autograd refuses to run a backward whose saved weight views were modified in
        # place after the forward — exactly what a pipelined step does).
        self.shadow = [nn.Parameter(v.detach(), requires_grad=True) for v in self.views[0]]
        opt_cls = getattr(torch.optim, optim_name)
        self.base_optimizer = opt_cls(self.shadow, **optimizer_args)
        # every slot starts as version 0 (reference: initialize_queue clones the same weights)
        self._queue: List[int] = list(range(self.num_versions))   # slots, oldest first
        self._versions_in_slot = [0] * self.num_versions
        # forwards always run on the NEWEST slot; new weights are written into the evicted
        # oldest one, so the storage a forward used stays intact for V-1 further steps
        self._point_to(self._queue[-1])

    # ------------------------------------------------------------------ torch.optim passthroughs
    def __getattr__(self, key):
        return getattr(self.base_optimizer, key)

    @property
    def param_groups(self):
        return self.base_optimizer.param_groups

    def zero_grad(self, set_to_none: bool = False):
        if self.batch_counter % self.update_interval == 0:
            for p in self.params:
                if p.grad is not None:
                    if set_to_none:
                        p.grad = None
                    else:
                        p.grad.detach_()
                        p.grad.zero_()

    # ------------------------------------------------------------------ versions
    def _point_to(self, slot: int) -> None:
        for p, v in zip(self.params, self.views[slot]):
            p.data = v
        self._current_slot = slot

    def load_old_params(self) -> None:
        """Weights the OLDEST in-flight forward saw (head of the ring)."""
        if self.num_versions > 1:
            self._point_to(self._queue[0])
            self.current_version = Version(self._versions_in_slot[self._queue[0]])

    def load_new_params(self) -> None:
        if self.num_versions > 1:
            self._point_to(self._queue[-1])
            self.current_version = self.latest_version

    def get_params(self, clone: bool = True):
        src = self.views[self._queue[-1]]
        return ([v.clone() for v in src] if clone else src), self.latest_version

    # ------------------------------------------------------------------ step
    def step(self, closure=None):
        self.batch_counter += 1
        if self.batch_counter % self.update_interval != 0:
            # macrobatch: keep accumulating; still rotate "old" so backward keeps matching
            return None
        newest = self._queue[-1]
        scale = 1.0 / (self.loss_scale * self.update_interval)
        if self.num_versions == 1:
            target = newest
        else:
            # the evicted oldest slot receives the new weights: copy latest -> oldest, update
            # in place there (no in-flight backward references it any more)
            target = self._queue.pop(0)
            self.flat[target].copy_(self.flat[newest])
        for q, v, p_ in zip(self.shadow, self.views[target], self.params):
            q.data = v
            q.grad = None if p_.grad is None else (p_.grad if scale == 1.0 else p_.grad * scale)
        loss = self.base_optimizer.step(closure)
        self.latest_version = self.latest_version.incr()
        self._versions_in_slot[target] = self.latest_version.version
        if self.num_versions > 1:
            self._queue.append(target)
        self._point_to(target)
        self.current_version = self.latest_version
        return loss

    def state_dict(self):
        return {"optimizer": self.base_optimizer.state_dict(),
                "latest_version": self.latest_version.version}

    def load_state_dict(self, sd):
        self.base_optimizer.load_state_dict(sd["optimizer"])
        self.latest_version = Version(sd.get("latest_version", 0))


class SGDWithWeightStashing(OptimizerWithWeightStashing):
    """``pipedream-fork/runtime/sgd.py:8-24``."""

    def __init__(self, modules, master_parameters=None, model_parameters=None, loss_scale=1.0,
                 num_versions=1, lr=0.1, momentum=0.0, dampening=0, weight_decay=0,
                 nesterov=False, verbose_freq=0, macrobatch=False):
        super().__init__("SGD", modules, master_parameters, model_parameters, loss_scale,
                         num_versions, verbose_freq, macrobatch, lr=lr, momentum=momentum,
                         dampening=dampening, weight_decay=weight_decay, nesterov=nesterov)


class AdamWithWeightStashing(OptimizerWithWeightStashing):
    """``pipedream-fork/runtime/adam.py:8-22``."""

    def __init__(self, modules, master_parameters=None, model_parameters=None, loss_scale=1.0,
                 num_versions=1, lr=1e-3, betas=(0.9, 0.999), weight_decay=0, verbose_freq=0,
                 macrobatch=False):
        super().__init__("Adam", modules, master_parameters, model_parameters, loss_scale,
                         num_versions, verbose_freq, macrobatch, lr=lr, betas=betas,
                         weight_decay=weight_decay)
