"""Per-layer forward/backward timing, parameter and activation sizes — the measurements the
PipeDream partitioner and ``balance_by_time`` consume.

Reference: ``torchprofiler.Profiling`` monkey-patches every leaf ``forward`` with
``time.time()`` + ``torch.cuda.synchronize()`` and times backward between a backward
*pre*-hook and a backward hook, which needs a patched PyTorch autograd
(``pipedream-fork/profiler/torchmodules/torchprofiler/profiling.py:129-168``,
``pipedream-fork/pre_hook.patch``); every measurement serialises the GPU.

Here the model runs on the static engine, whose executor calls ``begin/end`` around each
op's forward and backward; on GPUs those record CUDA events on the compute stream (C++
``LayerTimer`` in the extension) and are resolved with ONE synchronisation after the step.
Times of a fused op are spread over the layers it covers in proportion to a simple cost
model, so the emitted profile stays per-layer like the reference's ``graph.txt``.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from ..engine import StageExecutor
from ..models.graph import LayerGraphModel


class HostTimer:
    """CPU fallback with the same begin/end/resolve interface as the C++ ``LayerTimer``."""

    def __init__(self):
        self.spans: List[List[float]] = []

    def begin(self, ident: int) -> int:
        self.spans.append([ident, time.perf_counter(), 0.0])
        return len(self.spans) - 1

    def end(self, h: int) -> None:
        self.spans[h][2] = time.perf_counter()

    def resolve(self) -> List[Tuple[int, float]]:
        out = [(int(i), (b - a) * 1e3) for i, a, b in self.spans]
        self.spans = []
        return out


def make_timer(device) -> object:
    if torch.device(device).type == "cuda":
        from .. import ops
        return ops.ext().LayerTimer()
    return HostTimer()


@dataclass
class LayerProfile:
    names: List[str] = field(default_factory=list)
    descs: List[str] = field(default_factory=list)
    forward_ms: List[float] = field(default_factory=list)
    backward_ms: List[float] = field(default_factory=list)
    activation_bytes: List[float] = field(default_factory=list)
    parameter_bytes: List[float] = field(default_factory=list)
    output_shapes: List[Tuple[int, ...]] = field(default_factory=list)
    input_ms: float = 0.0


def _weight(mod: nn.Module) -> float:
    """Relative cost of a layer inside a fused op (conv/linear dominate)."""
    if isinstance(mod, (nn.Conv2d, nn.Linear)):
        return 8.0
    if isinstance(mod, nn.BatchNorm2d):
        return 2.0
    return 1.0


def profile_layers(model: LayerGraphModel, sample: torch.Tensor, device=None,
                   dtype: Optional[torch.dtype] = None, trials: int = 3,
                   warmup: int = 1, tcgen05: bool = False) -> LayerProfile:
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    dt = dtype or (torch.bfloat16 if dev.type == "cuda" else torch.float32)
    ex = StageExecutor.from_model(model, device=dev, dtype=dt, tcgen05=tcgen05, first_stage=False)
    x = ex.prepare_input(sample)
    target = torch.zeros(sample.shape[0], dtype=torch.int64, device=dev)
    n_nodes = len(model.nodes)
    fwd = [0.0] * n_nodes
    bwd = [0.0] * n_nodes
    shapes: Dict[int, Tuple[int, ...]] = {}
    for it in range(warmup + trials):
        timed = it >= warmup
        ex.profiler = make_timer(dev) if timed else None
        outs, ctx = ex.forward([x], training=True, need_input_grad=True)
        if it == 0:
            # output shape of every op (fused ops expose only their last layer's output)
            vals = {}
        stats, d = ex.loss(outs[0], target)
        ex.backward(ctx, [d])
        if timed:
            for ident, ms in ex.profiler.resolve():
                op = ex.plan.ops[ident // 2]
                ids = op.node_ids
                w = [_weight(model.node_module(i)) for i in ids]
                tot = sum(w)
                for i, wi in zip(ids, w):
                    (bwd if ident % 2 else fwd)[i] += ms * wi / tot / trials
    ex.profiler = None
    # shapes / sizes from a meta dry run of the plain graph
    prof = LayerProfile()
    live_shapes = _node_output_shapes(model, sample.shape[0])
    for i, n in enumerate(model.nodes):
        mod = model.node_module(i)
        prof.names.append(n.name)
        prof.descs.append(repr(mod).replace("\n", ""))
        prof.forward_ms.append(fwd[i])
        prof.backward_ms.append(bwd[i])
        shp = live_shapes[n.name]
        prof.output_shapes.append(shp)
        numel = 1
        for s in shp:
            numel *= s
        prof.activation_bytes.append(float(numel * 4))        # fp32 bytes, as the reference
        prof.parameter_bytes.append(float(sum(p.numel() for p in mod.parameters()) * 4))
    return prof


def _node_output_shapes(model: LayerGraphModel, batch: int) -> Dict[str, Tuple[int, ...]]:
    from ..models.graph import INPUT
    shapes: Dict[str, Tuple[int, ...]] = {}
    with torch.no_grad():
        live = {INPUT: torch.empty((batch,) + tuple(model.input_shape), device="meta")}
        shapes[INPUT] = tuple(live[INPUT].shape)
        for i, n in enumerate(model.nodes):
            mod = model.node_module(i)
            state = {k: torch.empty_like(v, device="meta")
                     for k, v in list(mod.named_parameters()) + list(mod.named_buffers())}
            was = mod.training
            mod.eval()
            out = torch.func.functional_call(mod, state, tuple(live[k] for k in n.inputs))
            mod.train(was)
            live[n.name] = out
            shapes[n.name] = tuple(out.shape)
    return shapes
