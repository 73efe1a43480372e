"""1F1B pipeline runtime for ARBITRARY ``nn.Module`` stages with tuple interfaces (autograd
path) — what the reference's ``StageRuntime`` is for its generated ``StageK`` modules
(``pipedream-fork/runtime/runtime.py:44-701``), used here for GNMT
(``runtime/translation/main_with_runtime.py``).  The image models run on the engine-based
``StageRuntime`` (``runtime.py`` in this package) instead.

* one rank per stage (straight pipeline; replicate a whole pipeline for data parallelism),
* static message shapes (batches are padded to fixed lengths) over ``DistP2P`` with
  pre-posted receives — no shape messages, no helper threads,
* integer tensors (token ids, lengths, targets) ride along the activations forward and get no
  gradient message backward,
* weight stashing through ``OptimizerWithWeightStashing`` (flat version slots, pointer swaps).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from ...dist.transport import DistP2P, PairGroups, TensorSpec
from .optimizer import OptimizerWithWeightStashing
from .schedule import F as OP_F, one_f_one_b


def infer_specs(stages: Sequence[nn.Module], sample_inputs: Sequence[torch.Tensor],
                input_names: Sequence[Sequence[str]], output_names: Sequence[Sequence[str]]
                ) -> List[List[TensorSpec]]:
    """Dry-run the chain on CPU-sized samples: specs[k] = tensors entering stage k
    (specs[len(stages)] = final outputs)."""
    specs = []
    vals = list(sample_inputs)
    names = list(input_names[0])
    with torch.no_grad():
        for k, st in enumerate(stages):
            specs.append([TensorSpec(n, tuple(v.shape), v.dtype) for n, v in zip(names, vals)])
            was = st.training
            st.eval()
            outs = st(*vals)
            st.train(was)
            named = dict(zip(output_names[k], outs))
            if k + 1 < len(stages):
                names = list(input_names[k + 1])
                vals = [named[n] for n in names]
            else:
                specs.append([TensorSpec(n, tuple(v.shape), v.dtype)
                              for n, v in zip(output_names[k], outs)])
    return specs


class ModuleStageRuntime:
    def __init__(self, stage: nn.Module, stage_index: int, num_stages: int,
                 in_specs: Sequence[TensorSpec], out_specs: Sequence[TensorSpec], device,
                 optimizer: Optional[OptimizerWithWeightStashing] = None,
                 loss_fn: Optional[Callable] = None, ranks: Optional[Sequence[int]] = None,
                 groups: Optional[PairGroups] = None, dtype: Optional[torch.dtype] = None):
        self.stage, self.k, self.n = stage, stage_index, num_stages
        self.device = torch.device(device)
        self.in_specs, self.out_specs = list(in_specs), list(out_specs)
        self.opt, self.loss_fn = optimizer, loss_fn
        self.ranks = list(ranks) if ranks is not None else list(range(num_stages))
        self.is_first, self.is_last = stage_index == 0, stage_index == num_stages - 1
        self.prev = None if self.is_first else self.ranks[stage_index - 1]
        self.next = None if self.is_last else self.ranks[stage_index + 1]
        self.warmup = num_stages - stage_index - 1
        self.p2p = DistP2P(self.device, groups, ring=self.warmup + 3) if num_stages > 1 else None
        self.grad_in = [s for s in self.in_specs if s.dtype.is_floating_point]
        self.grad_out = [s for s in self.out_specs if s.dtype.is_floating_point]
        self._saved: List[Tuple[list, list, Optional[torch.Tensor]]] = []
        self._next_fwd = self._next_bwd = None
        self.losses: List[torch.Tensor] = []
        self.training = True
        self._fwd_left = self._bwd_left = 0

    # ------------------------------------------------------------------ receive pre-posting
    def _post_fwd(self):
        # never post more receives than messages the peer will send in this epoch
        if not self.is_first and self._fwd_left > 0:
            self._fwd_left -= 1
            self._next_fwd = self.p2p.post_recv(self.prev, self.in_specs, "fwd", ring=self.warmup + 3)

    def _post_bwd(self):
        if not self.is_last and self.grad_out and self._bwd_left > 0:
            self._bwd_left -= 1
            self._next_bwd = self.p2p.post_recv(self.next, self.grad_out, "bwd", ring=3)

    # ------------------------------------------------------------------ passes
    def run_forward(self, inputs: Optional[Sequence[torch.Tensor]] = None, target=None):
        if self.is_first:
            ins = [t.to(self.device, non_blocking=True) for t in inputs]
        else:
            if self._next_fwd is None:
                self._post_fwd()
            ins = [t for t in self._next_fwd.wait()]
            self._next_fwd = None
            self._post_fwd()
        if self.training:
            ins = [t.detach().requires_grad_(True) if t.is_floating_point() and not self.is_first
                   else t for t in ins]
        with torch.set_grad_enabled(self.training):
            outs = self.stage(*ins)
        outs = list(outs) if isinstance(outs, (tuple, list)) else [outs]
        loss = None
        if self.is_last:
            if self.loss_fn is not None and target is not None:
                with torch.set_grad_enabled(self.training):
                    loss = self.loss_fn(outs[0], target.to(self.device))
                self.losses.append(loss.detach())
        else:
            self.p2p.send(self.next, [o.detach() for o in outs], "fwd")
        if self.training:
            self._saved.append((ins, outs, loss))
        return outs, loss

    def run_backward(self):
        ins, outs, loss = self._saved.pop(0)
        if self.opt is not None:
            self.opt.load_old_params()
        if self.is_last:
            loss.backward()
        else:
            if self._next_bwd is None:
                self._post_bwd()
            grads = list(self._next_bwd.wait())
            self._next_bwd = None
            self._post_bwd()
            fouts = [o for o in outs if o.is_floating_point()]
            pairs = [(o, g) for o, g in zip(fouts, grads) if o.requires_grad]
            torch.autograd.backward([o for o, _ in pairs], [g for _, g in pairs])
        if not self.is_first and self.grad_in:
            gin = [t.grad if t.grad is not None else torch.zeros_like(t)
                   for t in ins if t.is_floating_point()]
            self.p2p.send(self.prev, gin, "bwd")
        if self.opt is not None:
            self.opt.load_new_params()
            self.opt.step()
            self.opt.zero_grad()

    # ------------------------------------------------------------------ epoch drivers
    def run_epoch(self, loader, num_iterations: int, on_step: Optional[Callable] = None):
        """``loader`` yields (inputs tuple, target); consumed by the first stage (inputs) and the
        last stage (target) — every stage iterates the same deterministic loader."""
        self.training = True
        self.stage.train()
        self._fwd_left = self._bwd_left = num_iterations
        it = iter(loader) if loader is not None else None
        for op, i in one_f_one_b(num_iterations, self.warmup):
            if op == OP_F:
                inputs = target = None
                if it is not None:
                    inputs, target = next(it)
                outs, loss = self.run_forward(inputs if self.is_first else None,
                                              target if self.is_last else None)
                if on_step is not None:
                    on_step(i, loss)
            else:
                self.run_backward()
        if self.p2p is not None:
            self.p2p.flush()

    @torch.no_grad()
    def run_eval(self, loader, num_iterations: int):
        self.training = False
        self.stage.eval()
        self._fwd_left, self._bwd_left = num_iterations, 0
        it = iter(loader) if loader is not None else None
        results = []
        for _ in range(num_iterations):
            inputs = target = None
            if it is not None:
                inputs, target = next(it)
            outs, loss = self.run_forward(inputs if self.is_first else None,
                                          target if self.is_last else None)
            if self.is_last:
                results.append((outs, loss))
        if self.p2p is not None:
            self.p2p.flush()
        self.training = True
        return results
