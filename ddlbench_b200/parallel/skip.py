"""torchgpipe's skip-connection API — ``skippable`` / ``stash`` / ``pop`` / ``Namespace`` — and the
adapter that turns an ``nn.Sequential`` written against it into a ``LayerGraphModel``.

The reference's pipeline-ready models are plain ``nn.Sequential``s whose residual connections
cross layer (and therefore partition) boundaries through torchgpipe "portals"
(``benchmark/imagenet/gpipemodels/resnet/block.py:7,31-50``: ``Identity`` stashes the block
input, ``Residual`` pops it eleven layers later).  This framework represents a model as an explicit
layer GRAPH instead (``models/graph.py``): a skip connection is just a second input edge, every
cut is legal and the tensors that cross it are computed (``LayerGraphModel.live_after``) — no
portals, no hidden state.  ``sequential_to_graph`` bridges the two: the generator protocol of
every skippable layer is driven once on a sample input to find out WHICH tensors it stashes and
pops, and the result is an ordinary layer graph that ``GPipe`` / the PipeDream runtime can
partition anywhere.

    @skippable(stash=['identity'])
    class Identity(nn.Module):
        def forward(self, x):
            yield stash('identity', x)
            return x

    @skippable(pop=['identity'])
    class Residual(nn.Module):
        def forward(self, x):
            identity = yield pop('identity')
            return x + identity

    ns = Namespace()
    block = nn.Sequential(Identity().isolate(ns), conv, bn, Residual().isolate(ns), relu)
    model = sequential_to_graph(block, sample=torch.randn(1, 64, 8, 8))
"""
from __future__ import annotations

import inspect
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from ..models.graph import INPUT, LayerGraphModel, LayerNode

__all__ = ["Namespace", "stash", "pop", "skippable", "Skippable", "sequential_to_graph", "verify_skippables"]


class Namespace:
    """Isolates the skip names of one block from identically named skips elsewhere
    (``torchgpipe.skip.Namespace``)."""
    _count = 0

    def __init__(self):
        Namespace._count += 1
        self.id = Namespace._count

    def __repr__(self):
        return "<Namespace %d>" % self.id


class stash:  # noqa: N801 - torchgpipe's spelling
    """``yield stash(name, tensor)`` inside a skippable ``forward``."""
    __slots__ = ("name", "tensor")

    def __init__(self, name: str, tensor: Optional[torch.Tensor]):
        self.name, self.tensor = name, tensor


class pop:  # noqa: N801
    """``tensor = yield pop(name)`` inside a skippable ``forward``."""
    __slots__ = ("name",)

    def __init__(self, name: str):
        self.name = name


class Skippable(nn.Module):
    """A module whose ``forward`` is a generator speaking the stash/pop protocol."""

    def __init__(self, module: nn.Module, stashable: Sequence[str], poppable: Sequence[str]):
        super().__init__()
        self.module = module
        self.stashable_names = tuple(stashable)
        self.poppable_names = tuple(poppable)
        self.namespaces: Dict[str, Optional[Namespace]] = {}

    def isolate(self, ns: Namespace, *, only: Optional[Iterable[str]] = None) -> "Skippable":
        names = set(self.stashable_names) | set(self.poppable_names)
        for n in (names if only is None else set(only)):
            if n not in names:
                raise ValueError("'%s' is not a skip name of %s" % (n, type(self.module).__name__))
            self.namespaces[n] = ns
        return self

    def key(self, name: str) -> Tuple[Optional[int], str]:
        ns = self.namespaces.get(name)
        return (ns.id if ns is not None else None, name)

    def dispatch(self, x, popped: Dict[str, torch.Tensor]):
        """Drive the generator: feed ``popped`` tensors, collect stashes.  -> (output, stashes)."""
        out = self.module(x)
        stashes: Dict[str, torch.Tensor] = {}
        if not inspect.isgenerator(out):
            return out, stashes
        gen = out
        try:
            op = next(gen)
            while True:
                if isinstance(op, stash):
                    if op.name not in self.stashable_names:
                        raise RuntimeError("'%s' has not been declared as stashable" % op.name)
                    stashes[op.name] = op.tensor
                    op = next(gen)
                elif isinstance(op, pop):
                    if op.name not in self.poppable_names:
                        raise RuntimeError("'%s' has not been declared as poppable" % op.name)
                    if op.name not in popped:
                        raise RuntimeError("'%s' popped before it was stashed" % op.name)
                    op = gen.send(popped[op.name])
                else:
                    raise TypeError("a skippable forward must yield stash(...) or pop(...), got %r" % (op,))
        except StopIteration as stop:
            return stop.value, stashes

    def forward(self, x, *popped):
        """Graph-node calling convention: positional popped tensors in ``poppable_names`` order;
        returns the output (stashes are resolved to graph edges by ``sequential_to_graph``)."""
        out, _ = self.dispatch(x, dict(zip(self.poppable_names, popped)))
        return out


def skippable(stash: Iterable[str] = (), pop: Iterable[str] = ()):  # noqa: A002 - torchgpipe's signature
    """Class decorator: ``@skippable(stash=['a'], pop=['b']) class M(nn.Module): ...`` — instances
    are ``Skippable`` wrappers with ``.isolate(ns)``."""
    stash_names, pop_names = tuple(stash), tuple(pop)

    def wrap(cls):
        def make(*args, **kwargs):
            return Skippable(cls(*args, **kwargs), stash_names, pop_names)
        make.__name__ = cls.__name__
        make.__doc__ = cls.__doc__
        make.wrapped = cls
        return make
    return wrap


def _flatten(seq: nn.Sequential, prefix: str = "") -> List[Tuple[str, nn.Module]]:
    out = []
    for name, child in seq.named_children():
        full = prefix + name
        if isinstance(child, nn.Sequential):
            out += _flatten(child, full + ".")
        else:
            out.append((full, child))
    return out


def verify_skippables(seq: nn.Sequential) -> None:
    """Static check of torchgpipe's rules: every pop has an earlier stash of the same
    (namespace, name), nothing is stashed twice, nothing stays un-popped."""
    pending = {}
    for name, m in _flatten(seq):
        if not isinstance(m, Skippable):
            continue
        for n in m.poppable_names:
            if m.key(n) not in pending:
                raise TypeError("'%s' pops '%s' which no earlier layer stashes" % (name, n))
            del pending[m.key(n)]
        for n in m.stashable_names:
            if m.key(n) in pending:
                raise TypeError("'%s' stashes '%s' which is already stashed by '%s'" % (name, n, pending[m.key(n)]))
            pending[m.key(n)] = name
    if pending:
        raise TypeError("stashed but never popped: %s" % sorted(v for v in pending.values()))


class _StashTap(nn.Module):
    """Fallback for a stash that is neither the layer's input nor its output: re-runs the layer
    and returns the stashed intermediate (costs a second evaluation; deterministic layers only)."""

    def __init__(self, sk: Skippable, name: str):
        super().__init__()
        self.sk, self.name = sk, name

    def forward(self, x, *popped):
        _, stashes = self.sk.dispatch(x, dict(zip(self.sk.poppable_names, popped)))
        return stashes[self.name]


def sequential_to_graph(seq: nn.Sequential, sample: torch.Tensor, name: str = "sequential",
                        num_classes: Optional[int] = None) -> LayerGraphModel:
    """Convert a (possibly nested) ``nn.Sequential`` with skippable layers into a layer graph.
    ``sample``: one input batch; the model is run once on it (eval-free, no_grad) to learn which
    tensor every ``stash`` refers to."""
    verify_skippables(seq)
    nodes: List[LayerNode] = []
    edges: Dict[Tuple[Optional[int], str], str] = {}    # (namespace, skip name) -> producing graph value
    prev = INPUT
    x = sample
    _values.clear()
    _values[INPUT] = sample
    was_training = seq.training
    used = set()
    with torch.no_grad():
        for lname, m in _flatten(seq):
            node_name = lname if lname not in used and lname != INPUT else lname + "_"
            used.add(node_name)
            if isinstance(m, Skippable):
                pop_inputs = [edges.pop(m.key(n)) for n in m.poppable_names]
                popped_vals = {n: _values[e] for n, e in zip(m.poppable_names, pop_inputs)}
                out, stashes = m.dispatch(x, popped_vals)
                nodes.append(LayerNode(node_name, m, tuple([prev] + pop_inputs), lname.rsplit(".", 1)[0]))
                for sname, t in stashes.items():
                    if t is x:
                        edges[m.key(sname)] = prev                 # the layer's input: no new node
                    elif t is out:
                        edges[m.key(sname)] = node_name            # the layer's output
                    else:
                        tap = node_name + ".stash." + sname
                        nodes.append(LayerNode(tap, _StashTap(m, sname), tuple([prev] + pop_inputs),
                                               lname.rsplit(".", 1)[0]))
                        _values[tap] = t
                        edges[m.key(sname)] = tap
                _values[node_name] = out
                x = out
            else:
                x = m(x)
                nodes.append(LayerNode(node_name, m, (prev,), lname.rsplit(".", 1)[0]))
                _values[node_name] = x
            prev = node_name
    seq.train(was_training)
    _values.clear()
    ncls = num_classes if num_classes is not None else (int(x.shape[1]) if x.dim() == 2 else 0)
    return LayerGraphModel(nodes, name=name, input_shape=tuple(sample.shape[1:]), num_classes=ncls)


# values seen during the single tracing pass (name -> tensor); module-level scratch, cleared after use
_values: Dict[str, torch.Tensor] = {}
