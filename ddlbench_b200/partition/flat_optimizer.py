"""Partitioner for purely SEQUENTIAL models described by per-layer CSV profiles — the older
tool of the reference (``pipedream-fork/optimizer/optimizer.py:10-175`` + ``utils.py``, Python
2 there) with the same recurrence as the graph partitioner restricted to a chain, plus the
helpers of ``optimizer/scripts`` (CSV profile -> graph conversion, branch compression)."""
from __future__ import annotations

import argparse
import csv
import math
from typing import List, Tuple

from .graph import Graph, Node


def read_profile_csv(path: str) -> List[Tuple[str, float, float, float, float]]:
    """Rows: layer description, forward ms, backward ms, activation bytes, parameter bytes
    (header line optional)."""
    rows = []
    with open(path) as f:
        for r in csv.reader(f):
            if not r or not r[0].strip():
                continue
            try:
                rows.append((r[0], float(r[1]), float(r[2]), float(r[3]), float(r[4])))
            except (ValueError, IndexError):
                continue      # header
    return rows


def convert_profile_to_graph(rows) -> Graph:
    """``scripts/convert_profiles_to_graphs.py``: a CSV layer profile is a chain graph."""
    g = Graph()
    prev = None
    for i, (desc, f, b, a, p) in enumerate(rows):
        n = Node("node%d" % i, desc, f, b, a, p)
        g.add_node(n)
        if prev is not None:
            g.add_edge(prev, n)
        prev = n
    return g


def compress_graph_branches(g: Graph) -> Graph:
    """``scripts/compress_graph_branches.py``."""
    return g.compress_branches()


def optimize_chain(rows, num_machines: int, bandwidth: float, memory_size: float = math.inf,
                   straight_pipeline: bool = False) -> Tuple[float, List[int], List[int]]:
    """Returns (pipeline time per minibatch [s], stage end indices (exclusive), replication)."""
    n = len(rows)
    t = [(r[1] + r[2]) / 1000.0 for r in rows]
    act = [r[3] for r in rows]
    par = [r[4] for r in rows]
    pre_t, pre_p, pre_a = [0.0], [0.0], [0.0]
    for i in range(n):
        pre_t.append(pre_t[-1] + t[i])
        pre_p.append(pre_p[-1] + par[i])
        pre_a.append(pre_a[-1] + act[i])
    INF = float("inf")
    A = [[(INF, None, None)] * (num_machines + 1) for _ in range(n + 1)]
    A[0] = [(0.0, None, None)] * (num_machines + 1)
    for j in range(1, n + 1):
        for m in range(1, num_machines + 1):
            best = (INF, None, None)
            for k in range(0, j):
                for mp in range(1, (1 if straight_pipeline else m) + 1):
                    if k > 0 and m - mp < 1:
                        continue
                    prev = A[k][m - mp][0] if k > 0 else 0.0
                    if prev == INF:
                        continue
                    stage_t = pre_t[j] - pre_t[k]
                    stage_p = pre_p[j] - pre_p[k]
                    stash = (pre_a[j] - pre_a[k] + stage_p) * math.ceil((num_machines - m) / mp + 1e-12)
                    if stash > memory_size:
                        continue
                    stage = (stage_t + 4.0 * (mp - 1) * stage_p / (bandwidth * mp)) / mp
                    comm = 2.0 * act[k - 1] / (bandwidth * mp) if k > 0 else 0.0
                    val = max(prev, stage, comm)
                    if val < best[0]:
                        best = (val, k, mp)
            A[j][m] = best
    # back-track using all machines (or fewer if that is better)
    m = min(range(1, num_machines + 1), key=lambda mm: A[n][mm][0])
    ends, reps, j = [], [], n
    while j > 0:
        val, k, mp = A[j][m]
        ends.append(j)
        reps.append(mp)
        j, m = k, m - mp
    return A[n][max(1, sum(reps))][0] if reps else 0.0, ends[::-1], reps[::-1]


def cli(argv=None) -> int:
    ap = argparse.ArgumentParser(description="Sequential (CSV profile) PipeDream optimizer")
    ap.add_argument("-f", "--profile_filename", required=True)
    ap.add_argument("-n", "--num_machines", type=int, required=True)
    ap.add_argument("-b", "--network_bandwidth", type=float, default=1e9)
    ap.add_argument("-s", "--memory_size", type=float, default=math.inf)
    ap.add_argument("--straight_pipeline", action="store_true")
    a = ap.parse_args(argv)
    rows = read_profile_csv(a.profile_filename)
    t, ends, reps = optimize_chain(rows, a.num_machines, a.network_bandwidth, a.memory_size,
                                   a.straight_pipeline)
    print("(Split start, split end) / replication factor per stage:")
    prev = 0
    for e, r in zip(ends, reps):
        print((prev, e), r)
        prev = e
    print()
    print("Time per stage in pipeline:", t)
    return 0


if __name__ == "__main__":
    raise SystemExit(cli())
