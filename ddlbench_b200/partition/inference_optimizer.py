"""Partitioner variant for INFERENCE pipelines: minimise end-to-end latency subject to a
throughput constraint (one input every ``time_between_inputs`` seconds).

Same problem and cost model as ``pipedream-fork/optimizer/inference_optimizer_graph.py:14-205``:
only forward compute time counts; a stage replicated m' ways serves inputs round-robin, so its
*time per input* is ``stage_time / m'`` (must stay below the constraint) while its contribution
to the *latency* of one input is the full ``stage_time`` plus the activation transfers in and
out; ``B[i][m]`` = (time per input, latency) of the best way to run the states up to ``i`` on
``m+1`` workers, minimising latency among the feasible splits.
"""
from __future__ import annotations

import argparse
from typing import List, Optional, Tuple

from .graph import Graph


def main(num_machines: int, profile_filename: Optional[str], time_between_inputs: float,
         network_bandwidth: float, memory_size: float = 16e9, straight_pipeline: bool = False,
         use_memory_constraint: bool = False, use_fewer_machines: bool = False,
         activation_compression: bool = False, output_directory: Optional[str] = None,
         num_machines_in_first_level: Optional[int] = None, graph: Optional[Graph] = None,
         out=print) -> Tuple[Optional[float], Optional[float], List[int], List[int]]:
    """Returns (time per input, latency, split state indices, replication factors)."""
    if num_machines_in_first_level is not None and num_machines_in_first_level > num_machines:
        raise ValueError("num_machines_in_first_level has to be at most num_machines")
    gr = graph if graph is not None else Graph.from_str(open(profile_filename).read())
    adag = gr.antichain_dag()
    states = adag.topological_sort()
    idx = {s: i for i, s in enumerate(states)}
    _, _, _, order = gr._reach()
    for s in states:
        s.output_activation_size = sum(gr.nodes[a].activation_size for a in s.antichain)
        ct = at = pt = 0.0
        for i in gr._bits(gr._pred_mask(s.antichain)):
            p = gr.nodes[order[i]]
            ct += p.forward_compute_time / 1000.0
            at += p.activation_size
            pt += p.parameter_size
        s.compute_time, s.act_total, s.param_total = ct, at, pt
    S, M = len(states), num_machines
    NONE = (None, None, None, None)
    B = [[NONE] * M for _ in range(S)]
    for i, s in enumerate(states):
        for j in range(1 if straight_pipeline else M):
            if use_memory_constraint and s.act_total + s.param_total > memory_size:
                continue
            if num_machines_in_first_level is not None and j != num_machines_in_first_level - 1:
                continue
            if s.compute_time / (j + 1) < time_between_inputs:
                B[i][j] = (s.compute_time / (j + 1), s.compute_time, None, j + 1)
    preds = [[idx[p] for p in adag.predecessors(s.node_id)] for s in states]
    first = 1 if num_machines_in_first_level is None else num_machines_in_first_level
    for m in range(first, M):
        for i in range(1, S):
            best = B[i][m]
            if use_fewer_machines and m > 0 and (best[0] is None or (
                    B[i][m - 1][0] is not None and B[i][m - 1][0] < best[0])):
                best = B[i][m - 1]
            for j in preds[i]:
                for mp in range(1, 2 if straight_pipeline else m + 1):
                    prev = B[j][m - mp]
                    if prev[0] is None:
                        continue
                    in_t = states[j].output_activation_size / (network_bandwidth * mp)
                    out_t = (states[i].output_activation_size / (network_bandwidth * mp)
                             if i < S - 1 else None)
                    stage = states[i].compute_time - states[j].compute_time
                    stash = (states[i].act_total - states[j].act_total) + \
                            (states[i].param_total - states[j].param_total)
                    if use_memory_constraint and stash > memory_size:
                        continue
                    per_input = stage / mp
                    latency = prev[1] + stage
                    t = max(prev[0], per_input)
                    if not activation_compression:
                        t = max(t, in_t)
                        latency += in_t * mp
                        if out_t is not None:
                            t = max(t, out_t)
                            latency += out_t * mp
                    if t > time_between_inputs:
                        continue
                    if best[1] is None or best[1] > latency:
                        best = (t, latency, (j, m - mp), mp)
            B[i][m] = best
    meta = B[S - 1][M - 1]
    splits, reps = [], []
    nxt = meta[2]
    while nxt is not None:
        reps.append(meta[3])
        splits.append(nxt[0] + 1)
        meta = B[nxt[0]][nxt[1]]
        nxt = meta[2]
    if meta[3] is not None:
        reps.append(meta[3])
    splits.reverse()
    reps.reverse()
    final = B[S - 1][M - 1]
    out("(Split start, split end) / replication factor per stage:")
    prev = 0
    for k, sp in enumerate(splits + [S]):
        out((prev, sp), reps[k] if k < len(reps) else 1)
        prev = sp
    out()
    out("Time per input: %s, latency: %s" % (final[0], final[1]))
    if output_directory is not None and final[0] is not None:
        import os
        os.makedirs(output_directory, exist_ok=True)
        sid = 0
        prev = 0
        for sp in splits + [S]:
            for p in gr.all_predecessors(states[sp - 1].antichain):
                if p.stage_id is None:
                    p.set_stage_id(sid)
            sid += 1
        with open(os.path.join(output_directory, "gpus=%d.txt" % num_machines), "w") as f:
            f.write(str(gr))
    return final[0], final[1], splits, reps


def cli(argv=None) -> int:
    ap = argparse.ArgumentParser(description="PipeDream inference optimizer")
    ap.add_argument("-n", "--num_machines", type=int, required=True)
    ap.add_argument("-f", "--profile_filename", required=True)
    ap.add_argument("-t", "--time_between_inputs", type=float, required=True)
    ap.add_argument("-b", "--network_bandwidth", type=float, default=1e9)
    ap.add_argument("-s", "--memory_size", type=float, default=16e9)
    ap.add_argument("--straight_pipeline", action="store_true")
    ap.add_argument("--use_memory_constraint", action="store_true")
    ap.add_argument("--use_fewer_machines", action="store_true")
    ap.add_argument("--activation_compression", action="store_true")
    ap.add_argument("-o", "--output_directory", default=None)
    ap.add_argument("--num_machines_in_first_level", type=int, default=None)
    a = ap.parse_args(argv)
    main(a.num_machines, a.profile_filename, a.time_between_inputs, a.network_bandwidth,
         a.memory_size, a.straight_pipeline, a.use_memory_constraint, a.use_fewer_machines,
         a.activation_compression, a.output_directory, a.num_machines_in_first_level)
    return 0


if __name__ == "__main__":
    raise SystemExit(cli())
