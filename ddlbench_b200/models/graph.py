"""Layer-graph model representation.

Every network in the zoo is defined ONCE as an ordered list of named layer
nodes whose inputs are explicit edges (so residual / skip connections are
first-class).  From that single definition we derive the three "dialects" the
reference keeps as three separate hand-written copies of every network:

* a plain ``nn.Module`` (reference: ``benchmark/cifar10/pytorchcifargitmodels/*.py``,
  ``benchmark/mnist/models/*.py``, ``torchvision.models`` for ImageNet),
* a flat ``nn.Sequential`` in which *any* cut point is legal because the skip
  tensors travel with the activation tuple (reference:
  ``benchmark/*/gpipemodels/*/flatten_sequential.py`` + ``torchgpipe.skip``
  stash/pop "portals", ``benchmark/imagenet/gpipemodels/resnet/block.py:31-50``),
* a profiler graph (``graph.txt``) for the PipeDream partitioner (reference:
  ``pipedream-fork/profiler/torchmodules/torchgraph/graph_creator.py``; the
  reference needs a third copy of each model that avoids ``torch.nn.functional``
  so a tracer can see every op, ``README_PIPEDREAM.md:31-34``).
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import torch.nn as nn

INPUT = "input0"


class Add(nn.Module):
    """n-ary elementwise sum node (residual join)."""

    def forward(self, *xs):
        out = xs[0]
        for x in xs[1:]:
            out = out + x
        return out

    def __repr__(self):
        return "Add()"


class Flatten(nn.Module):
    """``x.view(x.size(0), -1)`` as a module (the reference models call ``view``)."""

    def forward(self, x):
        return x.reshape(x.size(0), -1)

    def __repr__(self):
        return "View(-1)"


@dataclass
class LayerNode:
    name: str
    op: nn.Module
    inputs: Tuple[str, ...]
    # tag used by block-level coarsening ("layer2.1") and by the fused-plan matcher
    block: str = ""


class GraphBuilder:
    """Incrementally builds the ordered node list of a network."""

    def __init__(self):
        self.nodes: List[LayerNode] = []
        self._names = set([INPUT])
        self._block = ""

    def block(self, name: str):
        self._block = name
        return self

    def add(self, name: str, op: nn.Module, *inputs: str) -> str:
        if name in self._names:
            raise ValueError(f"duplicate node name {name}")
        for i in inputs:
            if i not in self._names:
                raise ValueError(f"node {name}: unknown input {i}")
        self._names.add(name)
        self.nodes.append(LayerNode(name, op, tuple(inputs), self._block))
        return name


class LayerGraphModel(nn.Module):
    """Plain executable view of a layer graph.

    Nodes execute in definition order (which is a topological order by
    construction).  Tensors are dropped right after their last consumer so the
    eager path has the same live-set as a hand-written ``forward``.
    """

    def __init__(self, nodes: Sequence[LayerNode], name: str = "model",
                 input_shape: Tuple[int, ...] = (3, 224, 224), num_classes: int = 1000):
        super().__init__()
        self.arch = name
        self.input_shape = tuple(input_shape)
        self.num_classes = num_classes
        self.layers = nn.ModuleDict(OrderedDict((n.name.replace(".", "_"), n.op) for n in nodes))
        self._nodes = list(nodes)
        self._keys = [n.name.replace(".", "_") for n in nodes]
        self._last_use = self._compute_last_use(self._nodes)

    # ------------------------------------------------------------------ utils
    @staticmethod
    def _compute_last_use(nodes: Sequence[LayerNode]) -> Dict[str, int]:
        last: Dict[str, int] = {}
        for i, n in enumerate(nodes):
            for inp in n.inputs:
                last[inp] = i
        return last

    @property
    def nodes(self) -> List[LayerNode]:
        return self._nodes

    def node_module(self, i: int) -> nn.Module:
        return self.layers[self._keys[i]]

    # ---------------------------------------------------------------- forward
    def forward(self, x):
        live = {INPUT: x}
        out = x
        last = self._last_use
        for i, n in enumerate(self._nodes):
            mod = self.layers[self._keys[i]]
            out = mod(*[live[k] for k in n.inputs])
            live[n.name] = out
            for k in n.inputs:
                if last.get(k, -1) == i:
                    del live[k]
        return out

    # ------------------------------------------------------------- live sets
    def live_after(self, i: int) -> List[str]:
        """Names of tensors that must cross a cut placed right after node ``i``
        (main activation first).  This is what torchgpipe's stash/pop portals
        carry implicitly; here it is explicit and computed, so every cut is legal."""
        produced = {INPUT} | {n.name for n in self._nodes[: i + 1]}
        needed: List[str] = []
        for j in range(i + 1, len(self._nodes)):
            for inp in self._nodes[j].inputs:
                if inp in produced and inp not in needed:
                    needed.append(inp)
        main = self._nodes[i].name
        if main in needed:
            needed.remove(main)
            needed.insert(0, main)
        return needed

    def as_sequential(self) -> "nn.Sequential":
        """Flat sequential view: element ``i`` maps the live tuple before node
        ``i`` to the live tuple after it."""
        return nn.Sequential(OrderedDict(
            (self._keys[i], _SeqStep(self, i)) for i in range(len(self._nodes))))

    def segment(self, start: int, stop: int) -> "GraphSegment":
        return GraphSegment(self, start, stop)

    def block_boundaries(self) -> List[int]:
        """Indices ``i`` such that cutting after node ``i`` crosses no skip edge
        (live set is exactly one tensor)."""
        return [i for i in range(len(self._nodes) - 1) if len(self.live_after(i)) == 1]


class _SeqStep(nn.Module):
    """One node wrapped so it maps live-tuple → live-tuple."""

    def __init__(self, parent: LayerGraphModel, i: int):
        super().__init__()
        self.op = parent.node_module(i)
        node = parent.nodes[i]
        self.in_names = parent.live_after(i - 1) if i > 0 else [INPUT]
        self.out_names = parent.live_after(i) if i < len(parent.nodes) - 1 else [node.name]
        self.node_inputs = node.inputs
        self.node_name = node.name

    def forward(self, *xs):
        if len(xs) == 1 and isinstance(xs[0], (tuple, list)):
            xs = tuple(xs[0])
        live = dict(zip(self.in_names, xs))
        live[self.node_name] = self.op(*[live[k] for k in self.node_inputs])
        outs = tuple(live[k] for k in self.out_names)
        return outs[0] if len(outs) == 1 else outs


class GraphSegment(nn.Module):
    """Contiguous slice ``[start, stop)`` of a layer graph as one module: a
    pipeline *stage*.  Inputs/outputs are the live sets at the two cuts."""

    def __init__(self, parent: LayerGraphModel, start: int, stop: int):
        super().__init__()
        assert 0 <= start < stop <= len(parent.nodes)
        self.start, self.stop = start, stop
        self.in_names = parent.live_after(start - 1) if start > 0 else [INPUT]
        self.out_names = (parent.live_after(stop - 1) if stop < len(parent.nodes)
                          else [parent.nodes[stop - 1].name])
        self._nodes = parent.nodes[start:stop]
        self._keys = [n.name.replace(".", "_") for n in self._nodes]
        self.layers = nn.ModuleDict(OrderedDict(
            (k, parent.node_module(start + j)) for j, k in enumerate(self._keys)))
        last: Dict[str, int] = {}
        for j, n in enumerate(self._nodes):
            for inp in n.inputs:
                last[inp] = j
        self._last_use = last

    def forward(self, *xs):
        if len(xs) == 1 and isinstance(xs[0], (tuple, list)):
            xs = tuple(xs[0])
        live = dict(zip(self.in_names, xs))
        keep = set(self.out_names)
        for j, n in enumerate(self._nodes):
            live[n.name] = self.layers[self._keys[j]](*[live[k] for k in n.inputs])
            for k in n.inputs:
                if self._last_use.get(k) == j and k not in keep:
                    live.pop(k, None)
        outs = tuple(live[k] for k in self.out_names)
        return outs[0] if len(outs) == 1 else outs


def init_kaiming(model: nn.Module) -> None:
    """torchvision-style init (reference: ``benchmark/imagenet/gpipemodels/resnet/resnet.py:74-84``)."""
    for m in model.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, (nn.BatchNorm2d,)):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Linear):
            nn.init.normal_(m.weight, 0, 0.01)
            nn.init.zeros_(m.bias)
