"""Compile a layer graph (or a contiguous slice of one — a pipeline stage) into a list of
fused ops.

Pattern matching over the single model definition (``models/graph.py``):

    Conv2d(bias=False) -> BatchNorm2d [-> Add(other)] [-> ReLU/ReLU6]   => ConvBnAct
    Conv2d(bias)       [-> ReLU]                                         => ConvBiasAct
    Linear             [-> ReLU]                                         => LinearAct
    AdaptiveAvgPool2d(1) / AvgPool2d(full map) -> Flatten                => GlobalAvgPoolFlatten
    Add [-> ReLU]                                                        => AddAct
    anything else                                                        => ModuleOp (autograd fallback)

A node is only absorbed into its producer when the intermediate value has no other
consumer and does not cross the slice boundary.

Reference counterpart: none at run time (eager ``nn.Module`` calls, one library kernel per layer);
the closest artefact is the code generator's stage modules
(``pipedream-fork/optimizer/convert_graph_to_model.py``), which this plan replaces for execution —
a stage here is a slice ``[start, stop)`` of the same layer graph, compiled, not generated source.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Set, Tuple

import torch.nn as nn

from ..models.graph import Add, Flatten, INPUT, LayerGraphModel
from ..ops import ACT_NONE, ACT_RELU, ACT_RELU6
from . import ops as E


def _act_of(m: nn.Module) -> Optional[int]:
    if isinstance(m, nn.ReLU6):
        return ACT_RELU6
    if isinstance(m, nn.ReLU):
        return ACT_RELU
    return None


class Plan:
    def __init__(self, ops_: List[E.Op], in_names: List[str], out_names: List[str], arch: str,
                 input_shape: Tuple[int, ...]):
        self.ops = ops_
        self.in_names = in_names
        self.out_names = out_names
        self.arch = arch
        self.input_shape = input_shape
        for i, op in enumerate(self.ops):
            op.order = i

    def param_specs(self):
        specs = []
        for op in self.ops:
            specs.extend(op.param_specs())
        return specs

    def buffers(self):
        out = {}
        for op in self.ops:
            out.update(op.buffers())
        return out

    def summary(self) -> str:
        kinds: Dict[str, int] = {}
        for op in self.ops:
            kinds[op.kind] = kinds.get(op.kind, 0) + 1
        return ", ".join(f"{k}x{v}" for k, v in sorted(kinds.items()))


def compile_plan(model: LayerGraphModel, start: int = 0, stop: Optional[int] = None) -> Plan:
    nodes = model.nodes
    stop = len(nodes) if stop is None else stop
    key = lambda n: n.name.replace(".", "_")
    pname = lambda n: "layers." + key(n)            # state-dict prefix of a node's module
    in_names = model.live_after(start - 1) if start > 0 else [INPUT]
    out_names = model.live_after(stop - 1) if stop < len(nodes) else [nodes[stop - 1].name]
    boundary: Set[str] = set(out_names)

    consumers: Dict[str, List[int]] = {}
    for i, n in enumerate(nodes):
        for inp in n.inputs:
            consumers.setdefault(inp, []).append(i)

    def sole_next(i: int) -> Optional[int]:
        """Index of the single consumer of node i's output if it is inside the slice and the
        value does not leave the slice, else None."""
        name = nodes[i].name
        cs = consumers.get(name, [])
        if len(cs) != 1 or name in boundary:
            return None
        j = cs[0]
        return j if start <= j < stop else None

    used: Set[int] = set()
    out: List[E.Op] = []
    i = start
    while i < stop:
        if i in used:
            i += 1
            continue
        n = nodes[i]
        m = n.op
        before = set(used)
        used.add(i)
        if isinstance(m, nn.Conv2d):
            j = sole_next(i)
            if m.bias is None and j is not None and isinstance(nodes[j].op, nn.BatchNorm2d) \
                    and nodes[j].inputs == (n.name,) and nodes[j].op.track_running_stats:
                bn_i = j
                chain, last, act, res = [bn_i], bn_i, ACT_NONE, None
                k = sole_next(last)
                if k is not None and isinstance(nodes[k].op, Add) and len(nodes[k].inputs) == 2 \
                        and nodes[last].name in nodes[k].inputs:
                    other = [x for x in nodes[k].inputs if x != nodes[last].name]
                    # the residual operand must already exist when this op runs
                    produced = {INPUT} | {nodes[t].name for t in range(0, i)}
                    if len(other) == 1 and other[0] in produced:
                        res, last = other[0], k
                        chain.append(k)
                        k = sole_next(last)
                    else:
                        k = None
                if k is not None and _act_of(nodes[k].op) is not None and \
                        nodes[k].inputs == (nodes[last].name,):
                    act, last = _act_of(nodes[k].op), k
                    chain.append(k)
                used.update(chain)
                inputs = [n.inputs[0]] + ([res] if res is not None else [])
                out.append(E.ConvBnAct(n.name, inputs, nodes[last].name, m, nodes[bn_i].op, act,
                                       res is not None, pname(n), pname(nodes[bn_i])))
            else:
                act, last = ACT_NONE, i
                if j is not None and _act_of(nodes[j].op) is not None:
                    act, last = _act_of(nodes[j].op), j
                    used.add(j)
                out.append(E.ConvBiasAct(n.name, [n.inputs[0]], nodes[last].name, m, act, pname(n)))
        elif isinstance(m, nn.Linear):
            act, last = ACT_NONE, i
            j = sole_next(i)
            if j is not None and _act_of(nodes[j].op) is not None:
                act, last = _act_of(nodes[j].op), j
                used.add(j)
            out.append(E.LinearAct(n.name, [n.inputs[0]], nodes[last].name, m, act, pname(n)))
        elif isinstance(m, (nn.AdaptiveAvgPool2d, nn.AvgPool2d)):
            j = sole_next(i)
            is_global = isinstance(m, nn.AdaptiveAvgPool2d) and \
                tuple(E._pair(m.output_size)) == (1, 1)
            if j is not None and isinstance(nodes[j].op, Flatten) and \
                    (is_global or getattr(m, "_ddl_global", False)):
                used.add(j)
                out.append(E.GlobalAvgPoolFlatten(n.name, [n.inputs[0]], nodes[j].name))
            else:
                out.append(E.AvgPool(n.name, [n.inputs[0]], n.name, m))
        elif isinstance(m, nn.MaxPool2d):
            out.append(E.MaxPool(n.name, [n.inputs[0]], n.name, m))
        elif isinstance(m, Flatten):
            out.append(E.FlattenOp(n.name, [n.inputs[0]], n.name))
        elif isinstance(m, nn.Dropout):
            out.append(E.DropoutOp(n.name, [n.inputs[0]], n.name, m))
        elif isinstance(m, Add):
            act, last = ACT_NONE, i
            j = sole_next(i)
            if j is not None and _act_of(nodes[j].op) is not None:
                act, last = _act_of(nodes[j].op), j
                used.add(j)
            out.append(E.AddAct(n.name, list(n.inputs), nodes[last].name, act))
        elif isinstance(m, nn.BatchNorm2d) and m.track_running_stats and len(n.inputs) == 1:
            act, last = ACT_NONE, i
            j = sole_next(i)
            if j is not None and _act_of(nodes[j].op) is not None and \
                    nodes[j].inputs == (n.name,):
                act, last = _act_of(nodes[j].op), j
                used.add(j)
            out.append(E.BnAct(n.name, [n.inputs[0]], nodes[last].name, m, act, pname(n)))
        elif type(m).__name__ == "_ShiftCrop":
            out.append(E.ShiftCropOp(n.name, [n.inputs[0]], n.name))
        elif type(m).__name__ == "Concat" and hasattr(m, "dim"):
            out.append(E.ConcatOp(n.name, list(n.inputs), n.name, m.dim))
        elif _act_of(m) is not None and len(n.inputs) == 1:
            out.append(E.AddAct(n.name, list(n.inputs), n.name, _act_of(m)))
        else:
            out.append(E.ModuleOp(n.name, list(n.inputs), n.name, m, pname(n)))
        out[-1].node_ids = sorted(used - before)   # layer-graph nodes this op covers
        i += 1
    if E.BN_FOLD:
        _mark_bn_folds(out, set(out_names))
    return Plan(out, list(in_names), list(out_names), model.arch, model.input_shape)


def _mark_bn_folds(ops_: List[E.Op], boundary: Set[str]) -> int:
    """Defer the application of a conv+BN(+res)+act op into its FIRST consumer when that consumer
    is a 1x1 / stride-1 convolution taking the value as its convolution input (it then runs as
    the BN-fold GEMM).  Later consumers (skip connections) find the value already written."""
    marked = 0
    for i, p in enumerate(ops_):
        if p.kind != "conv_bn_act" or p.output in boundary:
            continue
        first = next((q for q in ops_[i + 1:] if p.output in q.inputs), None)
        if first is None or first.kind not in ("conv_bn_act", "conv_bias_act"):
            continue
        co, ci, kh, kw = first.wshape
        if first.inputs[0] != p.output or (kh, kw) != (1, 1) or first.stride != [1, 1] or \
                first.padding != [0, 0] or first.groups != 1 or ci % 8 or co % 8:
            continue
        p.defer_apply = True
        first.folds_input = True
        marked += 1
    return marked
