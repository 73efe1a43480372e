"""PipeDream hierarchical partitioner: profile graph -> pipeline stages + replication factors.

Same problem statement, cost model and stdout/file contracts as the reference tool DDLBench
drives (``pipedream-fork/optimizer/optimizer_graph_hierarchical.py``; invoked at
``run/run/run_template.sh:436-443`` and its stdout parsed at ``:445-462``):

* states = antichains of the profile graph in topological order (``partition.graph``);
* per hierarchy level, ``A[i][j][m]`` = best pipeline time of states (i, j] on m+1 workers,
  minimised over the last split point k and the replication m' of the last stage, with the
  data-parallel term ``4 (m'-1) params / (bw m')``, optional activation-transfer terms and
  the stashed-memory constraint ``ceil((M-(m+1))/m') (act+param) <= memory``;
* back-tracking prints ``(start, end) time replication`` per stage after a line containing
  "Split start, split end", a blank line ends the block, ``Level N`` banners separate levels;
* nodes are annotated with ``stage_id`` and written to ``gpus=<N>.txt`` (+ ``.dot``).

What is new: the O(S^2 M^2) recurrence runs in C++ (``partition_dp`` in the extension,
``csrc/bindings.cpp``) with a pure-Python fallback; on one NVSwitch domain the hierarchy is a
single level (``-n 8 -b 9e11``) — the reference models PCIe inside a node and Ethernet between
nodes (``run_template.sh:414-415``).
"""
from __future__ import annotations

import argparse
import math
import os
from collections import OrderedDict
from typing import List, Optional, Sequence

from .graph import Graph

NONE = -1.0


def _native():
    try:
        from .. import ops
        return ops.load_ext(build_if_missing=False).partition_dp
    except BaseException:  # noqa: BLE001
        return None


def compute_partitioning_py(compute, act, param, out_act, preds, num_machines, machines_within,
                            bandwidth, final_level, straight_pipeline, use_memory_constraint,
                            memory_size, use_fewer_machines, act_compression_ratio):
    """Pure-Python version of the recurrence; returns the same flat [R][S][M][4] table as the
    C++ ``partition_dp`` (time, split_k, split_m, replicas; -1 = None)."""
    R, S, Mn = len(compute), (len(compute[0]) if compute else 0), num_machines
    A = [NONE] * (R * S * Mn * 4)

    def at(i, j, m):
        return ((i * S + j) * Mn + m) * 4

    for i in range(R):
        for j in range(i, S):
            ct, a_s, p_s = compute[i][j], act[i][j], param[i][j]
            max_m = 1 if straight_pipeline else Mn
            for m in range(max_m):
                stashed = math.ceil((Mn - (m + 1)) / (m + 1)) * (a_s + p_s)
                if use_memory_constraint and stashed > memory_size:
                    continue
                dp_comm = (4.0 * m * p_s) / (bandwidth * (m + 1)) / machines_within
                if ct < 0:
                    continue
                o = at(i, j, m)
                A[o], A[o + 1], A[o + 2], A[o + 3] = (ct + dp_comm) / (m + 1), NONE, NONE, m + 1
    pred_sets = [set(p) for p in preds]
    max_i = 1 if final_level else R
    for i in range(max_i):
        for m in range(1, Mn):
            for j in range(i + 1, S):
                o = at(i, j, m)
                bt, bk, bm, br = A[o], A[o + 1], A[o + 2], A[o + 3]
                if use_fewer_machines and m > 0:
                    p = at(i, j, m - 1)
                    if bt < 0 or (A[p] >= 0 and A[p] < bt):
                        bt, bk, bm, br = A[p], A[p + 1], A[p + 2], A[p + 3]
                for k in preds[j]:
                    if i > 0 and k in pred_sets[i - 1]:
                        continue
                    max_mp = 2 if straight_pipeline else m + 1
                    for mp in range(1, max_mp):
                        in_t = (2.0 * out_act[k]) / (bandwidth * mp)
                        out_t = (2.0 * out_act[j]) / (bandwidth * mp) if j < S - 1 else NONE
                        if k + 1 >= R:
                            continue
                        last = compute[k + 1][j]
                        if last < 0:
                            continue
                        last_p = param[k + 1][j]
                        stashed = (act[k + 1][j] + last_p) * math.ceil((Mn - (m + 1)) / mp)
                        if use_memory_constraint and stashed > memory_size:
                            continue
                        last = (last + (4.0 * (mp - 1) * last_p) / (bandwidth * mp)) / mp
                        s = at(i, k, m - mp)
                        if A[s] < 0:
                            continue
                        t = max(A[s], last)
                        if act_compression_ratio > 0:
                            t = max(t, in_t / act_compression_ratio)
                            if out_t >= 0:
                                t = max(t, out_t / act_compression_ratio)
                        if bt < 0 or bt > t:
                            bt, bk, bm, br = t, float(k), float(m - mp), float(mp)
                A[o], A[o + 1], A[o + 2], A[o + 3] = bt, bk, bm, br
    return A


class _Table:
    """View of the flat DP table as A[i][j][m] -> (time | None, split | None, replicas)."""

    def __init__(self, flat, R, S, M):
        self.flat, self.R, self.S, self.M = flat, R, S, M

    def get(self, i, j, m):
        o = ((i * self.S + j) * self.M + m) * 4
        t, k, km, r = self.flat[o:o + 4]
        split = None if k < 0 else (int(k), int(km))
        return (None if t < 0 else t), split, (None if r < 0 else int(r))


def _analyze(T: _Table, states, start, end, bandwidth, num_machines, act_ratio,
             print_configuration, verbose, out=print):
    meta = T.get(start, end - 1, num_machines - 1)
    next_split = meta[1]
    remaining = num_machines
    splits, reps = [], []
    prev_split = end - 1
    while next_split is not None:
        used = meta[2]
        if verbose:
            out("-------------------------------------")
            out("Number of machines used: %d..." % used)
            out("Split between layers %d and %d..." % (next_split[0], next_split[0] + 1))
            out("Split before antichain %s..." % (states[next_split[0] + 1].antichain))
        splits.append(next_split[0] + 1)
        ctime = states[prev_split - 1].compute_time - states[next_split[0]].compute_time
        psize = states[prev_split - 1].param_total - states[next_split[0]].param_total
        dp_t = (4 * (used - 1) * psize) / (bandwidth * used)
        pp_in = (2.0 * states[next_split[0]].output_activation_size / used) / bandwidth
        pp_out = (2.0 * states[prev_split - 1].output_activation_size / used) / bandwidth
        if act_ratio is not None:
            pp_in /= act_ratio
            pp_out /= act_ratio
        else:
            pp_in = pp_out = 0.0
        if verbose:
            out(("Compute time = %f, Data-parallel communication time = %f, "
                 "Pipeline-parallel communication time = %f...") % (
                ctime / used, dp_t / used, max(pp_in, pp_out)))
        prev_split = splits[-1]
        meta = T.get(start, next_split[0], next_split[1])
        next_split = meta[1]
        reps.append(used)
        remaining -= used
    used = meta[2]
    remaining -= used
    if verbose:
        out("-------------------------------------")
        out("Number of machines used: %d..." % used)
        ctime = states[prev_split - 1].compute_time
        psize = states[prev_split - 1].param_total
        out("Compute time = %f, Data-parallel communication time = %f..." % (
            ctime / used, ((4 * (used - 1) * psize) / (bandwidth * used)) / used))
        out("-------------------------------------")
    if print_configuration:
        out("Number of machines in budget not used: %d..." % remaining)
        out()
        out("(Split start, split end) / compute time taken per stage "
            "/ replication factor per stage:")
    prev = start
    splits.reverse()
    splits.append(end)
    reps.append(used)
    reps.reverse()
    for i, sp in enumerate(splits):
        t = states[sp - 1].compute_time - (states[prev - 1].compute_time if prev > 0 else 0.0)
        if print_configuration:
            out((prev, sp), t, reps[i])
        prev = sp
    if print_configuration:
        out()
    return splits[:-1], reps


def main(all_num_machines: Sequence[int], profile_filename: str,
         network_bandwidths: Sequence[float], memory_size: float = 16e9,
         straight_pipeline: bool = False, use_memory_constraint: bool = False,
         use_fewer_machines: bool = False, activation_compression_ratio: Optional[float] = None,
         output_directory: Optional[str] = None, print_configuration: bool = True,
         verbose: bool = False, graph: Optional[Graph] = None, out=print,
         force_python: bool = False):
    """Returns (pipeline_time, data_parallel_time, stage replication factors, graph)."""
    gr = graph if graph is not None else Graph.from_str(open(profile_filename).read())
    # inputs always live on the first stage: detach them while optimising
    removed: "OrderedDict" = OrderedDict()
    for src in gr.sources():
        if src.node_desc.startswith("Input"):
            src.forward_compute_time = src.backward_compute_time = 0.0
            src.activation_size = src.parameter_size = 0.0
            removed[src] = list(gr.edges.get(src.node_id, []))
            gr.remove_node(src)
    for sink in gr.sinks():
        if sink.node_desc.startswith("__getitem__"):
            gr.remove_node(sink)

    adag = gr.antichain_dag()
    states = adag.topological_sort()
    if verbose:
        out("Total number of states: %d" % len(states))
    index = {s: i for i, s in enumerate(states)}
    _, _, _, order = gr._reach()
    for s in states:
        s.output_activation_size = sum(gr.nodes[a].activation_size for a in s.antichain)
        ct = at = pt = 0.0
        # fixed (topological) summation order: results are reproducible run to run — the
        # reference sums over a Python set of nodes, whose order follows object addresses
        for i in gr._bits(gr._pred_mask(s.antichain)):
            p = gr.nodes[order[i]]
            ct += (p.forward_compute_time + p.backward_compute_time) / 1000.0
            at += p.activation_size
            pt += p.parameter_size
        s.compute_time, s.act_total, s.param_total = ct, at, pt
    S = len(states)
    out_act = [s.output_activation_size for s in states]
    preds = [[index[p] for p in adag.predecessors(s.node_id)] for s in states]

    def table(attr):
        rows = [[getattr(s, attr) for s in states]]
        for i in range(1, S + 1):
            base = getattr(states[i - 1], attr)
            rows.append([getattr(states[j], attr) - base if j > i - 1 else NONE for j in range(S)])
        return rows

    compute, act, param = table("compute_time"), table("act_total"), table("param_total")
    dp = None if force_python else _native()
    tables: List[_Table] = []
    within = 1
    ratio = activation_compression_ratio if activation_compression_ratio is not None else -1.0
    for level, (nm, bw) in enumerate(zip(all_num_machines, network_bandwidths)):
        out("Solving optimization problem with %d machines with inter-machine bandwidth of "
            "%.2f GB/s" % (nm, bw / 10 ** 9))
        final = level == len(network_bandwidths) - 1
        args = (compute, act, param, out_act, preds, nm, within, float(bw), final,
                straight_pipeline, use_memory_constraint, float(memory_size), use_fewer_machines,
                ratio)
        flat = dp(*args) if dp is not None else compute_partitioning_py(*args)
        T = _Table(flat, S + 1, S, nm)
        within = nm
        for i in range(S + 1):
            for j in range(S):
                t = T.get(i, j, nm - 1)[0]
                compute[i][j] = NONE if t is None else t
        tables.append(T)

    splits = [(0, S)]
    all_reps: List[int] = []
    for lvl in range(len(tables) - 1, -1, -1):
        out("======================================")
        out("Level %d" % (lvl + 1))
        out("======================================")
        new_splits, stage_id, level_reps = [], 0, []
        for (start, end) in splits:
            part, reps = _analyze(tables[lvl], states, start, end, network_bandwidths[lvl],
                                  all_num_machines[lvl], activation_compression_ratio,
                                  print_configuration, verbose, out)
            level_reps.extend(reps)
            sp = start
            for split in part:
                new_splits.append((sp, split))
                if lvl == 0:
                    for p in gr.all_predecessors(states[split - 1].antichain):
                        if p.stage_id is None:
                            p.set_stage_id(stage_id)
                sp = split
                stage_id += 1
            new_splits.append((sp, end))
            if lvl == 0:
                for p in gr.all_predecessors(states[end - 1].antichain):
                    if p.stage_id is None:
                        p.set_stage_id(stage_id)
            stage_id += 1
        out("Total number of stages: %d" % stage_id)
        splits = new_splits
        if lvl == 0:
            all_reps = level_reps
    for src, outs in removed.items():
        src.stage_id = 0
        for o in outs:
            gr.add_edge(src, o)

    total_machines = 1
    for nm in all_num_machines:
        total_machines *= nm
    if output_directory is not None:
        os.makedirs(output_directory, exist_ok=True)
        gr.to_dot(os.path.join(output_directory, "gpus=%d" % total_machines))
        with open(os.path.join(output_directory, "gpus=%d.txt" % total_machines), "w") as f:
            f.write(str(gr))

    total_time = states[-1].compute_time
    total_param = states[-1].param_total
    dp_time, within = total_time, 1
    for nm, bw in zip(all_num_machines, network_bandwidths):
        comm = ((4 * (nm - 1) * total_param) / (bw * nm)) / within
        dp_time = (dp_time + comm) / nm
        within = nm
    pp_time = tables[-1].get(0, S - 1, all_num_machines[-1] - 1)[0]
    if verbose:
        out()
        out("Time taken by single-stage pipeline:", total_time)
        out("Time per stage in pipeline:", pp_time)
        out("Throughput increase (compared to single machine):", total_time / pp_time)
        dp_str = ",".join(str(e) for e in all_num_machines)
        out(("[Note that single-machine and (%s)-machine DP might not fit given memory "
             "constraints]") % dp_str)
        out("Throughput increase of (%s)-machine DP compared to single machine:" % dp_str,
            total_time / dp_time)
        out("Throughput increase (compared to (%s)-machine DP):" % dp_str, dp_time / pp_time)
    return pp_time, dp_time, all_reps, gr


def stage_to_num_ranks_map(reps: Sequence[int]) -> str:
    """``"0:3,1:1"`` string consumed by the code generator (``--stage_to_num_ranks_map``);
    the reference builds it in bash from the optimizer's stdout
    (``run/run/run_template.sh:445-498``)."""
    return ",".join("%d:%d" % (i, r) for i, r in enumerate(reps))


def cli(argv=None) -> int:
    ap = argparse.ArgumentParser(description="PipeDream optimizer for replicated settings")
    ap.add_argument("-n", "--all_num_machines", nargs="+", type=int, required=True)
    ap.add_argument("-f", "--profile_filename", required=True)
    ap.add_argument("-b", "--network_bandwidths", type=float, nargs="+", default=[1e9])
    ap.add_argument("-s", "--memory_size", type=float, default=16e9)
    ap.add_argument("--straight_pipeline", action="store_true")
    ap.add_argument("-o", "--output_directory", default=None)
    ap.add_argument("--use_memory_constraint", action="store_true")
    ap.add_argument("--use_fewer_machines", action="store_true")
    ap.add_argument("--activation_compression_ratio", default=None, type=float)
    ap.add_argument("--python", action="store_true", help="force the pure-Python recurrence")
    a = ap.parse_args(argv)
    assert len(a.all_num_machines) == len(a.network_bandwidths)
    main(a.all_num_machines, a.profile_filename, a.network_bandwidths, a.memory_size,
         a.straight_pipeline, a.use_memory_constraint, a.use_fewer_machines,
         a.activation_compression_ratio, a.output_directory, verbose=True,
         force_python=a.python)
    return 0


if __name__ == "__main__":
    raise SystemExit(cli())
