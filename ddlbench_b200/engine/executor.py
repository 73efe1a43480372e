"""Static executor: runs a compiled ``Plan`` forward and backward without an autograd tape.

One ``StageExecutor`` is one model (single GPU / data parallel) or one pipeline stage
(GPipe / PipeDream).  It owns the flat ``ParamStore`` and the BatchNorm buffers, supports
several in-flight micro/minibatches (each forward returns a context that its backward
consumes), gradient accumulation across micro-batches, forward/backward on a chosen weight
VERSION (weight stashing), and fires ``on_bucket_ready(i)`` during backward the moment the
last gradient of flat-buffer bucket ``i`` has been written — the hook the data-parallel layer
uses to launch the fused all-reduce+SGD kernel for that bucket on the communication stream.

Reference counterpart: ``torch.autograd`` running the stage's ``nn.Module`` (forward in
``pipedream-fork/runtime/runtime.py::_run_forward``, ``torch.autograd.backward`` in
``::_run_backward``; the benchmark scripts' ``loss.backward()``), Horovod's per-parameter
gradient hooks (``imagenet_horovod.py:135-139``) and the patched backward pre-hook
(``pre_hook.patch``) the upstream profiler needs — all replaced by explicit per-op backward
functions and the two hook points above.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from .. import ops
from ..models.graph import LayerGraphModel
from . import ops as E
from .plan import Plan, compile_plan
from .store import ParamStore


class FwdContext:
    __slots__ = ("saved", "version", "inputs", "outputs", "need_input_grad")

    def __init__(self, n_ops: int, version: int):
        self.saved: List[Dict[str, object]] = [dict() for _ in range(n_ops)]
        self.version = version
        self.inputs: List[torch.Tensor] = []
        self.outputs: List[torch.Tensor] = []
        self.need_input_grad = False


class StageExecutor:
    def __init__(self, plan: Plan, device="cpu", dtype: torch.dtype = torch.float32,
                 channels_last: Optional[bool] = None, versions: int = 1, world: int = 1,
                 rank: int = 0, shard_optimizer: bool = False, allocator=None,
                 bucket_elems: int = 8 << 20, lr: float = 0.1, momentum: float = 0.9,
                 weight_decay: float = 0.0, first_stage: bool = True, tcgen05: bool = False):
        self.plan = plan
        self.device = torch.device(device)
        self.dtype = dtype
        self.channels_last = (self.device.type == "cuda") if channels_last is None else channels_last
        self.first_stage = first_stage
        self.store = ParamStore(plan.param_specs(), self.device, dtype, versions=versions,
                                bucket_elems=bucket_elems, world=world, rank=rank,
                                shard_optimizer=shard_optimizer, allocator=allocator, lr=lr,
                                momentum=momentum, weight_decay=weight_decay)
        self.buffers = {k: v.to(self.device) for k, v in plan.buffers().items()}
        for op in plan.ops:
            op.to_device(self.device)
        self.tcgen05 = tcgen05
        self.on_bucket_ready: Optional[Callable[[int], None]] = None
        self._consumers: Dict[str, int] = {}
        for op in plan.ops:
            for i in op.inputs:
                self._consumers[i] = self._consumers.get(i, 0) + 1
        # how many parameters of each bucket are produced by each op (for bucket triggers)
        lay = self.store.layout
        self._bucket_sizes = [len(b.params) for b in lay.buckets]
        self._last_op_of_value: Dict[str, int] = {}
        for idx, op in enumerate(plan.ops):
            for i in op.inputs:
                self._last_op_of_value[i] = idx
        self.xent_ws: Optional[ops.XentWorkspace] = None
        # optional per-op timer with begin(id) -> handle / end(handle) (utils.layer_profiler);
        # ids are 2*op_index for forward, 2*op_index+1 for backward
        self.profiler = None

    # ------------------------------------------------------------------ helpers
    @classmethod
    def from_model(cls, model: LayerGraphModel, start: int = 0, stop: Optional[int] = None,
                   **kw) -> "StageExecutor":
        plan = compile_plan(model, start, stop)
        kw.setdefault("first_stage", start == 0)
        return cls(plan, **kw)

    def prepare_input(self, x: torch.Tensor) -> torch.Tensor:
        x = x.to(self.device, non_blocking=True)
        if x.is_floating_point() and x.dtype != self.dtype:
            x = x.to(self.dtype)
        if self.channels_last and x.dim() == 4:
            x = x.contiguous(memory_format=torch.channels_last)
        return x

    # ------------------------------------------------------------------ forward
    def forward(self, inputs: Sequence[torch.Tensor], training: bool = True,
                version: Optional[int] = None,
                need_input_grad: Optional[bool] = None) -> Tuple[List[torch.Tensor], FwdContext]:
        plan = self.plan
        version = self.store.latest if version is None else version
        P = self.store.weights(version)
        B = self.buffers
        ctx = FwdContext(len(plan.ops), version)
        ctx.need_input_grad = (not self.first_stage) if need_input_grad is None else need_input_grad
        vals: Dict[str, torch.Tensor] = dict(zip(plan.in_names, inputs))
        keep = set(plan.out_names)
        E.TCGEN05["conv1x1"] = E.TCGEN05["linear"] = self.tcgen05
        prof = self.profiler
        for idx, op in enumerate(plan.ops):
            xs = [vals[i] for i in op.inputs]
            if E.PENDING_FOLD and not getattr(op, "folds_input", False):
                # a deferred BatchNorm reached an op that cannot fold it: apply it the plain way
                for x_ in xs:
                    fold = E.PENDING_FOLD.pop(id(x_), None)
                    if fold is not None:
                        fold.materialize()
            if prof is not None:
                h = prof.begin(2 * idx)
            vals[op.output] = op.forward(P, B, xs, ctx.saved[idx], training)
            if prof is not None:
                prof.end(h)
            for i in op.inputs:
                if self._last_op_of_value.get(i) == idx and i not in keep:
                    vals.pop(i, None)
        outs = [vals[o] for o in plan.out_names]
        if E.PENDING_FOLD:                   # nothing may leave the stage unwritten
            for o in outs:
                fold = E.PENDING_FOLD.pop(id(o), None)
                if fold is not None:
                    fold.materialize()
        if training:
            ctx.outputs = outs
        return outs, ctx

    # ------------------------------------------------------------------ backward
    def backward(self, ctx: FwdContext, grad_outputs: Sequence[Optional[torch.Tensor]],
                 accumulate: bool = False, final: bool = True) -> List[Optional[torch.Tensor]]:
        """Back-propagate ``grad_outputs`` (one per plan output).  ``accumulate``: add into the
        flat gradient buffer (micro-batches after the first).  ``final``: last micro-batch of
        the step, so bucket-ready triggers fire."""
        plan = self.plan
        P = self.store.weights(ctx.version)
        G = self.store.g_views
        grads: Dict[str, List[torch.Tensor]] = {}
        for name, g in zip(plan.out_names, grad_outputs):
            if g is not None:
                grads.setdefault(name, []).append(g)
        E.TCGEN05["conv1x1"] = E.TCGEN05["linear"] = self.tcgen05
        lay = self.store.layout
        remaining = list(self._bucket_sizes)
        in_set = set(plan.in_names)
        side = self._wgrad_stream()
        E.WGRAD_OVERLAP["stream"] = side

        def join_side():
            if side is not None and E.WGRAD_OVERLAP["keep"]:
                torch.cuda.current_stream(self.device).wait_stream(side)
                E.WGRAD_OVERLAP["keep"].clear()

        for idx in range(len(plan.ops) - 1, -1, -1):
            op = plan.ops[idx]
            dys = grads.pop(op.output, None)
            if dys is None:
                if op.params and not accumulate:
                    for p in op.params:   # no gradient reaches this op: its grads are zero
                        G[p].zero_()
            else:
                need = [(i not in in_set) or ctx.need_input_grad for i in op.inputs]
                if self.profiler is not None:
                    h = self.profiler.begin(2 * idx + 1)
                dxs = op.backward(P, G, dys, ctx.saved[idx], accumulate, need)
                if self.profiler is not None:
                    self.profiler.end(h)
                for i, dx in zip(op.inputs, dxs):
                    if dx is not None:
                        grads.setdefault(i, []).append(dx)
            ctx.saved[idx] = {}
            if final and self.on_bucket_ready is not None and op.params:
                for p in op.params:
                    b = lay.bucket_of[p]
                    remaining[b] -= 1
                    if remaining[b] == 0:
                        join_side()            # the bucket's weight gradients must have landed
                        self.on_bucket_ready(b)
        join_side()
        E.WGRAD_OVERLAP["stream"] = None
        out: List[Optional[torch.Tensor]] = []
        for name in plan.in_names:
            gl = grads.get(name)
            if not gl or not ctx.need_input_grad:
                out.append(None)
            else:
                g = gl[0]
                for extra in gl[1:]:
                    g = g + extra
                out.append(g)
        return out

    def _wgrad_stream(self):
        """Side stream for weight-gradient kernels (GPU only; ``overlap_wgrad`` attribute)."""
        if self.device.type != "cuda" or not getattr(self, "overlap_wgrad", E.OVERLAP_WGRAD_DEFAULT):
            return None
        if getattr(self, "_wg_stream", None) is None:
            self._wg_stream = torch.cuda.Stream(device=self.device)
        return self._wg_stream

    # ------------------------------------------------------------------ loss (last stage)
    def loss(self, logits: torch.Tensor, target: torch.Tensor, smoothing: float = 0.0,
             grad_scale: float = 1.0, want_grad: bool = True):
        """Fused softmax-cross-entropy on the stage output.  Returns (stats[3], dlogits)."""
        if logits.is_cuda and self.xent_ws is None:
            self.xent_ws = ops.XentWorkspace(max(logits.shape[0], 1024), logits.device)
        return ops.xent(logits, target, smoothing, grad_scale, want_grad, self.xent_ws)

    # ------------------------------------------------------------------ state
    def state_dict(self) -> Dict[str, torch.Tensor]:
        sd = self.store.state_dict()
        sd.update({k: v.detach().cpu().clone() for k, v in self.buffers.items()})
        return sd

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        self.store.load_state_dict({k: sd[k] for k in self.store.layout.entries})
        for k in self.buffers:
            if k in sd:
                self.buffers[k].copy_(sd[k].to(self.device))

    def num_params(self) -> int:
        return sum(e.numel for e in self.store.layout.entries.values())
