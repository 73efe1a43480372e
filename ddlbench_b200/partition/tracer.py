"""Trace an arbitrary ``nn.Module`` into the profile-graph IR.

Reference: ``torchgraph.GraphCreator`` replaces every leaf ``forward`` with a wrapper that
passes ``TensorWrapper`` objects overloading ``size/view/__add__/__mul__/__getitem__/
transpose`` and ``cat`` to record ``Size/View/Add/Mul/__getitem__/Transpose/Concat`` nodes
(``pipedream-fork/profiler/torchmodules/torchgraph/graph_creator.py:55-288``) — which is why
profiled models may not use ``torch.nn.functional`` (``README_PIPEDREAM.md:31-34``).

Here ``torch.fx`` symbolic tracing does the same job for any traceable module: ``call_module``
nodes become layer nodes (description = ``repr(module)``), ``+``/``*``/``cat``/``view``/
``flatten``/``size``/``getitem``/``transpose`` become the reference's graph-only op names.  The
zoo of this framework does not need tracing (its models ARE layer graphs); this is for
user-supplied modules.  ``to_layer_graph`` additionally converts a traced module into a
``LayerGraphModel`` when it only uses modules, adds and flattens.
"""
from __future__ import annotations

import operator
from typing import Dict

import torch
import torch.fx as fx
import torch.nn as nn

from ..models.graph import Add, Flatten, GraphBuilder, INPUT, LayerGraphModel
from .graph import Graph, Node

_FUNC_NAMES = {operator.add: "Add", torch.add: "Add", operator.iadd: "Add",
               operator.mul: "Mul", torch.mul: "Mul", torch.cat: "Concat",
               torch.flatten: "View(-1)", operator.getitem: "__getitem__",
               torch.transpose: "Transpose"}
_METHOD_NAMES = {"view": "View", "reshape": "View", "flatten": "View(-1)", "size": "Size",
                 "transpose": "Transpose", "add": "Add", "mul": "Mul", "contiguous": None}


def trace_to_graph(module: nn.Module, concrete_args=None) -> Graph:
    gm = fx.symbolic_trace(module, concrete_args=concrete_args)
    g = Graph()
    ids: Dict[fx.Node, Node] = {}
    alias: Dict[fx.Node, fx.Node] = {}
    n_in = 0
    k = 0
    for node in gm.graph.nodes:
        if node.op == "placeholder":
            nd = Node("node%d" % k, "Input%d" % n_in)
            n_in += 1
        elif node.op == "call_module":
            sub = gm.get_submodule(node.target)
            nd = Node("node%d" % k, repr(sub).replace("\n", ""),
                      parameter_size=float(sum(p.numel() for p in sub.parameters()) * 4))
        elif node.op == "call_function":
            name = _FUNC_NAMES.get(node.target, getattr(node.target, "__name__", str(node.target)))
            if node.target is operator.getitem:
                name = "__getitem__(%s)" % (node.args[1],)
            nd = Node("node%d" % k, name if "(" in name else name + "()")
        elif node.op == "call_method":
            name = _METHOD_NAMES.get(node.target, node.target)
            if name is None:          # no-op for the IR (e.g. .contiguous())
                alias[node] = node.args[0]
                continue
            if name == "View" and len(node.args) > 1:
                rest = [a for a in node.args[1:]]
                name = "View(-1)" if rest[-1] == -1 and len(rest) == 2 else "View(%s)" % (
                    ", ".join(str(a) for a in rest if not isinstance(a, fx.Node)) or "-1")
            elif name == "Size":
                name = "Size(%s)" % (node.args[1] if len(node.args) > 1 else "")
            nd = Node("node%d" % k, name if "(" in name else name + "()")
        elif node.op == "output":
            continue
        else:                          # get_attr
            nd = Node("node%d" % k, "Attr(%s)" % node.target)
        k += 1
        g.add_node(nd)
        ids[node] = nd
        for a in node.all_input_nodes:
            while a in alias:
                a = alias[a]
            if a in ids:
                g.add_edge(ids[a], nd)
    return g


def to_layer_graph(module: nn.Module, input_shape, num_classes: int = 0,
                   name: str = "traced") -> LayerGraphModel:
    """Convert a traceable module built from leaf modules, ``+`` and ``flatten/view(N,-1)``
    into a ``LayerGraphModel`` (so it can run on the engine, be partitioned and pipelined)."""
    gm = fx.symbolic_trace(module)
    b = GraphBuilder()
    names: Dict[fx.Node, str] = {}
    for node in gm.graph.nodes:
        ins = [names[a] for a in node.all_input_nodes if a in names]
        if node.op == "placeholder":
            names[node] = INPUT
        elif node.op == "call_module":
            names[node] = b.add(str(node.target), gm.get_submodule(node.target), *ins)
        elif node.op == "call_function" and node.target in (operator.add, torch.add, operator.iadd):
            names[node] = b.add(node.name, Add(), *ins)
        elif (node.op == "call_function" and node.target is torch.flatten) or \
                (node.op == "call_method" and node.target in ("flatten", "view", "reshape")):
            names[node] = b.add(node.name, Flatten(), ins[0])
        elif node.op == "call_method" and node.target in ("contiguous",):
            names[node] = ins[0]
        elif node.op == "call_method" and node.target == "size":
            continue
        elif node.op == "output":
            continue
        else:
            raise NotImplementedError("op %s/%s has no layer-graph equivalent" % (node.op, node.target))
    return LayerGraphModel(b.nodes, name=name, input_shape=tuple(input_shape), num_classes=num_classes)
