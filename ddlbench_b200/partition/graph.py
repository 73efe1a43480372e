"""Profile-graph IR shared by the profiler, the partitioner and the stage code generator.

File format (``graph.txt``) is byte-compatible with the reference's
(``pipedream-fork/graph/graph.py:451-480,635-663``): one line per node

    <id> -- <module repr> -- forward_compute_time=F, backward_compute_time=B, activation_size=A, parameter_size=P[ -- stage_id=S]

(times in ms, sizes in bytes; ``activation_size`` may be a ``[a; b; c]`` list) followed by
tab-indented edge lines ``\\t<src id> -- <dst id>``.

The implementation is new: nodes are indexed, reachability is kept as Python-int bitsets
computed in one topological sweep, and the antichain state space is enumerated iteratively —
the reference memoises recursive ``predecessors()`` / ``successors()`` per node
(``graph.py:319-348``) and hits recursion and time limits on deep residual networks
(ResNet-152 is disabled for PipeDream there, ``run/run/run.sh:56-62``).
"""
from __future__ import annotations

from collections import OrderedDict, deque
from typing import Dict, Iterable, List, Optional, Sequence, Set, Tuple


class Node:
    __slots__ = ("node_id", "node_desc", "forward_compute_time", "backward_compute_time",
                 "activation_size", "parameter_size", "stage_id", "depth", "height")

    def __init__(self, node_id: str, node_desc: str = "", forward_compute_time: float = 0.0,
                 backward_compute_time: float = 0.0, activation_size=0.0,
                 parameter_size: float = 0.0, stage_id: Optional[int] = None):
        self.node_id = node_id
        self.node_desc = node_desc
        self.forward_compute_time = forward_compute_time
        self.backward_compute_time = backward_compute_time
        self.activation_size = activation_size
        self.parameter_size = parameter_size
        self.stage_id = stage_id
        self.depth = None
        self.height = None

    def set_stage_id(self, stage_id: int) -> None:
        self.stage_id = stage_id

    def __str__(self) -> str:
        stage = " -- stage_id=%d" % self.stage_id if self.stage_id is not None else ""
        desc = self.node_desc.replace("\n", "")
        act = ("%s" % self.activation_size).replace(", ", "; ")
        return ("%s -- %s -- forward_compute_time=%.3f, backward_compute_time=%.3f, "
                "activation_size=%s, parameter_size=%.3f%s" % (
                    self.node_id, desc, self.forward_compute_time, self.backward_compute_time,
                    act, self.parameter_size, stage))

    @staticmethod
    def from_str(line: str) -> "Node":
        tok = line.strip().split(" -- ")
        node_id, desc, meta = tok[0], tok[1], tok[2]
        stage = int(tok[3].split("=")[1]) if len(tok) > 3 else None
        f, b, a, p = meta.split(", ")
        a = a.split("=", 1)[1]
        if "[" in a:
            act = sum(float(x) for x in a.strip("[]").split("; "))
        else:
            act = float(a)
        return Node(node_id, desc, float(f.split("=")[1]), float(b.split("=")[1]), act,
                    float(p.split("=")[1]), stage)

    def __repr__(self) -> str:
        return "Node(%s)" % self.node_id


class AntichainNode(Node):
    __slots__ = ("antichain", "output_activation_size", "compute_time", "act_total",
                 "param_total")

    def __init__(self, node_id: str, antichain: List[str], node_desc: str = ""):
        super().__init__(node_id, node_desc)
        self.antichain = antichain
        self.output_activation_size = 0.0
        self.compute_time = 0.0
        self.act_total = 0.0
        self.param_total = 0.0

    def __str__(self) -> str:
        return "%s -- %s" % (self.node_id, self.antichain)


class Graph:
    def __init__(self, node: Optional[Node] = None):
        self.nodes: "OrderedDict[str, Node]" = OrderedDict()
        self.edges: Dict[str, List[Node]] = {}
        self.in_edges: Dict[str, List[Node]] = {}
        self._reach_cache = None
        self._antichain_dag: Optional["Graph"] = None
        if node is not None:
            self.nodes[node.node_id] = node

    # ------------------------------------------------------------------ construction
    def add_node(self, node: Node) -> None:
        self.nodes[node.node_id] = node
        self._invalidate()

    def add_edge(self, a: Node, b: Node) -> None:
        self.nodes.setdefault(a.node_id, a)
        self.nodes.setdefault(b.node_id, b)
        self.edges.setdefault(a.node_id, []).append(b)
        self.in_edges.setdefault(b.node_id, []).append(a)
        self._invalidate()

    def remove_edge(self, a: Node, b: Node) -> None:
        self.edges[a.node_id].remove(b)
        self.in_edges[b.node_id].remove(a)
        self._invalidate()

    def remove_node(self, node: Node) -> None:
        nid = node.node_id
        del self.nodes[nid]
        for out in self.edges.pop(nid, []):
            self.in_edges[out.node_id].remove(node)
        for inn in self.in_edges.pop(nid, []):
            self.edges[inn.node_id].remove(node)
        self._invalidate()

    def _invalidate(self) -> None:
        self._reach_cache = None
        self._antichain_dag = None

    def reset(self) -> None:
        self._invalidate()

    def copy(self) -> "Graph":
        g = Graph()
        for n in self.nodes.values():
            g.nodes[n.node_id] = n
        for src, outs in self.edges.items():
            for o in outs:
                g.add_edge(self.nodes[src], o)
        return g

    def sources(self) -> List[Node]:
        return [n for nid, n in self.nodes.items() if not self.in_edges.get(nid)]

    def sinks(self) -> List[Node]:
        return [n for nid, n in self.nodes.items() if not self.edges.get(nid)]

    # ------------------------------------------------------------------ (de)serialisation
    def __str__(self) -> str:
        lines = [str(n) for n in self.nodes.values()]
        for n in self.nodes.values():
            for i in self.in_edges.get(n.node_id, []):
                lines.append("\t%s -- %s" % (i.node_id, n.node_id))
        return "\n".join(lines)

    @staticmethod
    def from_str(text: str) -> "Graph":
        g = Graph()
        for line in text.strip("\n").split("\n"):
            if not line.strip():
                continue
            if not line.startswith("\t"):
                n = Node.from_str(line)
                g.nodes[n.node_id] = n
            else:
                a, b = line.strip().split(" -- ")
                g.edges.setdefault(a, []).append(g.nodes[b])
                g.in_edges.setdefault(b, []).append(g.nodes[a])
        return g

    # ------------------------------------------------------------------ orderings
    def _heights(self) -> Dict[str, int]:
        order = self._kahn()
        h: Dict[str, int] = {}
        for nid in reversed(order):
            outs = self.edges.get(nid, [])
            h[nid] = 1 + max((h[o.node_id] for o in outs), default=0)
        return h

    def populate_heights(self) -> None:
        for nid, v in self._heights().items():
            self.nodes[nid].height = v

    def populate_depths(self) -> None:
        d: Dict[str, int] = {}
        for nid in self._kahn():
            ins = self.in_edges.get(nid, [])
            d[nid] = 1 + max((d[i.node_id] for i in ins), default=0)
            self.nodes[nid].depth = d[nid]

    def _kahn(self) -> List[str]:
        indeg = {nid: len(self.in_edges.get(nid, [])) for nid in self.nodes}
        q = deque(nid for nid in self.nodes if indeg[nid] == 0)
        out: List[str] = []
        while q:
            nid = q.popleft()
            out.append(nid)
            for o in self.edges.get(nid, []):
                indeg[o.node_id] -= 1
                if indeg[o.node_id] == 0:
                    q.append(o.node_id)
        if len(out) != len(self.nodes):
            raise Exception("Graph has a cycle")
        return out

    def topological_sort(self) -> List[Node]:
        """Deterministic DFS post-order (ties broken on node description, then height) —
        the order the partitioner's state indices and the generated code rely on."""
        # ties: node description, then the `height` attribute if populate_heights() was called
        # (same key as the reference's sort, pipedream-fork/graph/graph.py:311-312)
        heights = {nid: (n.height if n.height is not None else 0) for nid, n in self.nodes.items()}
        marked: Set[str] = set()
        result: List[str] = []
        roots = sorted(self.nodes.values(), key=lambda n: n.node_desc)
        for root in roots:
            if root.node_id in marked:
                continue
            stack: List[Tuple[str, int, List[Node]]] = []
            outs = sorted(self.edges.get(root.node_id, []),
                          key=lambda n: (n.node_desc, heights[n.node_id]))
            stack.append((root.node_id, 0, outs))
            on_stack = {root.node_id}
            while stack:
                nid, idx, outs = stack.pop()
                if idx < len(outs):
                    stack.append((nid, idx + 1, outs))
                    child = outs[idx].node_id
                    if child in marked:
                        continue
                    if child in on_stack:
                        raise Exception("Graph has a cycle")
                    couts = sorted(self.edges.get(child, []),
                                   key=lambda n: (n.node_desc, heights[n.node_id]))
                    stack.append((child, 0, couts))
                    on_stack.add(child)
                else:
                    marked.add(nid)
                    on_stack.discard(nid)
                    result.append(nid)
        result.reverse()
        return [self.nodes[nid] for nid in result]

    # ------------------------------------------------------------------ reachability
    def _reach(self):
        """(index of node id, predecessor bitsets, successor bitsets, id list)."""
        if self._reach_cache is not None:
            return self._reach_cache
        order = self._kahn()
        idx = {nid: i for i, nid in enumerate(order)}
        pred = [0] * len(order)
        succ = [0] * len(order)
        for nid in order:
            i = idx[nid]
            m = 0
            for p in self.in_edges.get(nid, []):
                j = idx[p.node_id]
                m |= pred[j] | (1 << j)
            pred[i] = m
        for nid in reversed(order):
            i = idx[nid]
            m = 0
            for s in self.edges.get(nid, []):
                j = idx[s.node_id]
                m |= succ[j] | (1 << j)
            succ[i] = m
        self._reach_cache = (idx, pred, succ, order)
        return self._reach_cache

    @staticmethod
    def _bits(mask: int) -> Iterable[int]:
        while mask:
            low = mask & -mask
            yield low.bit_length() - 1
            mask ^= low

    def predecessors(self, node_id: str) -> Set[Node]:
        idx, pred, _, order = self._reach()
        return {self.nodes[order[i]] for i in self._bits(pred[idx[node_id]])}

    def successors(self, node_id: str) -> Set[Node]:
        idx, _, succ, order = self._reach()
        return {self.nodes[order[i]] for i in self._bits(succ[idx[node_id]])}

    def all_predecessors(self, antichain: Sequence[str]) -> Set[Node]:
        idx, pred, _, order = self._reach()
        m = 0
        for nid in antichain:
            m |= pred[idx[nid]] | (1 << idx[nid])
        return {self.nodes[order[i]] for i in self._bits(m)}

    def _pred_mask(self, antichain: Sequence[str]) -> int:
        idx, pred, _, _ = self._reach()
        m = 0
        for nid in antichain:
            m |= pred[idx[nid]] | (1 << idx[nid])
        return m

    # ------------------------------------------------------------------ antichains
    def augment_antichain(self, antichain: Sequence[str]) -> List[str]:
        """Add every predecessor that still has an edge leaving the predecessor set of the
        antichain node it precedes (its output must also cross the cut)."""
        idx, pred, _, order = self._reach()
        extra: List[str] = []
        seen = set(antichain)
        for a in antichain:
            pm = pred[idx[a]]
            for i in self._bits(pm):
                nid = order[i]
                if nid in seen:
                    continue
                for o in self.edges.get(nid, []):
                    j = idx[o.node_id]
                    if not (pm >> j) & 1 and o.node_id != a:
                        extra.append(nid)
                        seen.add(nid)
                        break
        return extra + list(antichain)

    def deaugment_augmented_antichain(self, augmented: Sequence[str]) -> List[str]:
        idx, _, succ, _ = self._reach()
        amask = 0
        for a in augmented:
            amask |= 1 << idx[a]
        out: List[str] = []
        for a in augmented:
            if succ[idx[a]] & amask:
                continue          # an element that precedes another one is redundant
            if a not in out:
                out.append(a)
        return out

    def is_next_antichain(self, augmented: Sequence[str], new_node: str) -> bool:
        idx, _, succ, _ = self._reach()
        amask = 0
        for a in augmented:
            amask |= 1 << idx[a]
        return not (succ[idx[new_node]] & amask)

    def next_antichains(self, antichain: Sequence[str]) -> List[List[str]]:
        aset = set(antichain)
        augmented = self.augment_antichain(list(antichain))
        out: List[List[str]] = []
        for a in augmented:
            for nxt in self.edges.get(a, []):
                if nxt.node_id in aset:
                    continue
                if self.is_next_antichain(augmented, nxt.node_id):
                    repl = [x if x != a else nxt.node_id for x in augmented]
                    out.append(self.deaugment_augmented_antichain(repl))
        return out

    def antichain_dag(self) -> "Graph":
        """DAG whose nodes are the valid cuts ("antichains") of this graph; a path through it
        is a sequence of ever-growing prefixes, which lets a 1-D dynamic program partition a
        branching model.  Breadth-first, iterative."""
        if self._antichain_dag is not None:
            return self._antichain_dag
        dag = Graph()
        start = [self.sources()[0].node_id]
        key = lambda a: tuple(sorted(a))
        src = AntichainNode("antichain_0", self.augment_antichain(start))
        dag.nodes[src.node_id] = src
        dag.source = src
        mapping = {key(start): src}
        expanded: Set[Tuple[str, ...]] = set()
        q = deque([start])
        n_id = 0
        while q:
            a = q.popleft()
            k = key(a)
            if k in expanded:
                continue
            expanded.add(k)
            for nxt in self.next_antichains(a):
                nk = key(nxt)
                if nk not in mapping:
                    n_id += 1
                    mapping[nk] = AntichainNode("antichain_%d" % n_id, self.augment_antichain(nxt))
                dag.add_edge(mapping[k], mapping[nk])
                if nk not in expanded:
                    q.append(nxt)
        self._antichain_dag = dag
        return dag

    # ------------------------------------------------------------------ partitioning helpers
    def partition_graph(self) -> List["Graph"]:
        stage_ids = sorted({n.stage_id for n in self.nodes.values()},
                           key=lambda s: (s is None, s))
        if len(stage_ids) == 1:
            return [self.copy()]
        subs = []
        for sid in stage_ids:
            sub = Graph()
            for nid, n in self.nodes.items():
                if n.stage_id != sid:
                    continue
                sub.add_node(n)
                for o in self.edges.get(nid, []):
                    if o.stage_id == sid:
                        sub.add_edge(n, o)
            subs.append(sub)
        return subs

    def is_series_parallel(self) -> bool:
        """True iff repeated series/parallel reductions collapse the graph to one edge."""
        multi: Dict[Tuple[str, str], int] = {}
        succs: Dict[str, Set[str]] = {nid: set() for nid in self.nodes}
        preds: Dict[str, Set[str]] = {nid: set() for nid in self.nodes}
        for a, outs in self.edges.items():
            for o in outs:
                succs[a].add(o.node_id)
                preds[o.node_id].add(a)
        srcs = [n for n in self.nodes if not preds[n]]
        snks = [n for n in self.nodes if not succs[n]]
        if len(srcs) != 1 or len(snks) != 1:
            return False
        changed = True
        alive = set(self.nodes)
        while changed:
            changed = False
            for n in list(alive):
                if n in (srcs[0], snks[0]):
                    continue
                if len(preds[n]) == 1 and len(succs[n]) == 1:
                    (p,), (s,) = tuple(preds[n]), tuple(succs[n])
                    succs[p].discard(n)
                    preds[s].discard(n)
                    succs[p].add(s)
                    preds[s].add(p)
                    alive.discard(n)
                    changed = True
        return alive == {srcs[0], snks[0]} or len(alive) == 1

    def check_fidelity(self, other: "Graph") -> None:
        """Same multiset of node totals (used after compression / regeneration)."""
        for attr in ("forward_compute_time", "backward_compute_time", "parameter_size"):
            a = sum(getattr(n, attr) for n in self.nodes.values())
            b = sum(getattr(n, attr) for n in other.nodes.values())
            assert abs(a - b) <= 1e-6 * max(1.0, abs(a)), (attr, a, b)

    def check_isomorphism(self, other: "Graph") -> None:
        a, b = self.topological_sort(), other.topological_sort()
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert x.node_desc == y.node_desc, (x.node_desc, y.node_desc)
            assert len(self.edges.get(x.node_id, [])) == len(other.edges.get(y.node_id, []))
            assert len(self.in_edges.get(x.node_id, [])) == len(other.in_edges.get(y.node_id, []))

    def compress_branches(self) -> "Graph":
        """Collapse every maximal single-entry/single-exit linear chain into one node (sums
        of times/params, activation of the last node) — the coarsening used for
        Inception/NASNet-style graphs (reference ``graph.py:139-227``)."""
        g = Graph()
        order = self._kahn()
        head: Dict[str, str] = {}
        merged: Dict[str, Node] = {}
        for nid in order:
            n = self.nodes[nid]
            ins = self.in_edges.get(nid, [])
            if len(ins) == 1 and len(self.edges.get(ins[0].node_id, [])) == 1 and \
                    len(self.edges.get(nid, [])) <= 1 and ins[0].node_id in head:
                h = head[ins[0].node_id]
                m = merged[h]
                m.node_desc = m.node_desc + " + " + n.node_desc if len(m.node_desc) < 400 else m.node_desc
                m.forward_compute_time += n.forward_compute_time
                m.backward_compute_time += n.backward_compute_time
                m.parameter_size += n.parameter_size
                m.activation_size = n.activation_size
                head[nid] = h
            else:
                head[nid] = nid
                merged[nid] = Node(nid, n.node_desc, n.forward_compute_time,
                                   n.backward_compute_time, n.activation_size, n.parameter_size,
                                   n.stage_id)
        for nid in order:
            g.nodes.setdefault(head[nid], merged[head[nid]])
        seen = set()
        for a, outs in self.edges.items():
            for o in outs:
                ha, hb = head[a], head[o.node_id]
                if ha != hb and (ha, hb) not in seen:
                    seen.add((ha, hb))
                    g.add_edge(merged[ha], merged[hb])
        return g

    # ------------------------------------------------------------------ rendering
    def to_dot_source(self) -> str:
        colors = ["lightblue", "green", "grey", "firebrick1", "gold", "chocolate1", "beige"]
        lines = ["digraph G {"]
        for n in self.nodes.values():
            color = colors[n.stage_id % len(colors)] if n.stage_id is not None else "white"
            label = "%s\\n[fwd %.3f, bwd %.3f, act %s, param %.0f]" % (
                n.node_desc.replace('"', "'")[:60], n.forward_compute_time,
                n.backward_compute_time, n.activation_size, n.parameter_size)
            lines.append('  "%s" [label="%s", style=filled, fillcolor=%s];' % (n.node_id, label, color))
        for a, outs in self.edges.items():
            for o in outs:
                lines.append('  "%s" -> "%s";' % (a, o.node_id))
        lines.append("}")
        return "\n".join(lines)

    def to_dot(self, path: str) -> None:
        """Writes ``<path>.dot`` (and renders a PDF when graphviz is importable)."""
        src = self.to_dot_source()
        with open(path + ".dot", "w") as f:
            f.write(src)
        try:
            import graphviz  # type: ignore
            graphviz.Source(src).render(path, format="pdf", cleanup=True)
        except Exception:  # noqa: BLE001 - graphviz is optional
            pass
