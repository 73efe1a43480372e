"""Discrete-event simulation of the pipeline schedules (1F1B with stage replication, GPipe
fill-drain): given per-stage forward / backward times and boundary transfer times it returns the
steady-state throughput, the bubble fraction per rank and a Chrome-trace Gantt chart.

Purpose: check a partition BEFORE spending GPUs on it — the partitioner's objective ("time per
stage in pipeline") is a steady-state bound that ignores warm-up, the 1F1B ordering and
round-robin replica routing; the simulator replays the very schedule the runtime executes
(``parallel/pipedream/schedule.py``) with those costs.  The reference has no such tool (its
README tells the user to run the job and read ``slurm.out``).

    python -m ddlbench_b200.partition.simulate --fwd 2,2,2,2 --bwd 4,4,4,4 --comm 0.1 --replicas 1,1,1,1 -n 64
"""
from __future__ import annotations

import argparse
import json
from typing import Dict, List, Optional, Sequence, Tuple

from ..parallel.pipedream.schedule import F as OP_F, fill_drain, one_f_one_b


def simulate_1f1b(fwd: Sequence[float], bwd: Sequence[float], replicas: Optional[Sequence[int]] = None,
                  comm: float = 0.0, minibatches: int = 64, extra_warmup: int = 0) -> Dict[str, object]:
    """``fwd[s]`` / ``bwd[s]``: ms per minibatch on ONE replica of stage s.  Minibatch m is served
    by replica ``m % replicas[s]`` of every stage (the runtime's round-robin routing)."""
    S = len(fwd)
    replicas = list(replicas) if replicas is not None else [1] * S
    # per (stage, replica): the ordered op list of the runtime
    # warm-up forwards per replica: the runtime's formula (parallel/pipedream/config.py)
    num_ranks = sum(replicas)
    warm = []
    for s in range(S):
        w = num_ranks - 1 - sum(replicas[:s])
        warm.append(max(0, w // replicas[s]) + extra_warmup)
    f_done: Dict[Tuple[int, int], float] = {}      # (stage, minibatch) -> time forward finished
    b_done: Dict[Tuple[int, int], float] = {}
    free = {(s, r): 0.0 for s in range(S) for r in range(replicas[s])}
    busy = {k: 0.0 for k in free}
    spans: List[Tuple[int, int, str, int, float, float]] = []
    # every replica executes its own 1F1B sequence over the minibatches routed to it
    seqs = {}
    for s in range(S):
        for r in range(replicas[s]):
            mine = [m for m in range(minibatches) if m % replicas[s] == r]
            seqs[(s, r)] = [(op, mine[i]) for op, i in one_f_one_b(len(mine), min(warm[s], len(mine)))]
    pos = {k: 0 for k in seqs}
    remaining = sum(len(v) for v in seqs.values())
    while remaining:
        progressed = False
        for key, seq in seqs.items():
            while pos[key] < len(seq):
                op, m = seq[pos[key]]
                s, r = key
                if op == OP_F:
                    dep = 0.0 if s == 0 else f_done.get((s - 1, m))
                    if dep is None:
                        break
                    start = max(free[key], dep + (comm if s > 0 else 0.0))
                    end = start + fwd[s]
                    f_done[(s, m)] = end
                else:
                    dep = f_done.get((s, m)) if s == S - 1 else b_done.get((s + 1, m))
                    if dep is None:
                        break
                    start = max(free[key], dep + (comm if s < S - 1 else 0.0))
                    end = start + bwd[s]
                    b_done[(s, m)] = end
                free[key] = end
                busy[key] += end - start
                spans.append((s, r, op, m, start, end))
                pos[key] += 1
                remaining -= 1
                progressed = True
        if not progressed:
            raise RuntimeError("schedule deadlocked in simulation (inconsistent replication factors?)")
    makespan = max(free.values())
    # steady state: slope of the first stage's backward completions over the MIDDLE half of the
    # run (warm-up at the start and the drain at the end are both faster than the steady state)
    last = sorted(b_done[(0, m)] for m in range(minibatches))
    lo, hi = minibatches // 4, (3 * minibatches) // 4
    steady = (last[hi] - last[lo]) / (hi - lo) if hi > lo else makespan / max(minibatches, 1)
    return {"makespan_ms": makespan, "steady_ms_per_minibatch": steady,
            "throughput_per_s": 1e3 / steady if steady > 0 else float("inf"),
            "bubble_fraction": {"%d.%d" % k: 1.0 - busy[k] / makespan for k in busy},
            "bound_ms": max((f + b) / r for f, b, r in zip(fwd, bwd, replicas)), "spans": spans}


def simulate_gpipe(fwd: Sequence[float], bwd: Sequence[float], microbatches: int, comm: float = 0.0,
                   recompute: bool = True) -> Dict[str, object]:
    """One GPipe mini-batch: all micro-batch forwards, then all backwards (with re-computation of
    every partition but the last when ``recompute``), then the synchronous update."""
    S = len(fwd)
    free = [0.0] * S
    f_done: Dict[Tuple[int, int], float] = {}
    b_done: Dict[Tuple[int, int], float] = {}
    spans = []
    for op, m in fill_drain(microbatches):
        order = range(S) if op == OP_F else range(S - 1, -1, -1)
        for s in order:
            if op == OP_F:
                dep = 0.0 if s == 0 else f_done[(s - 1, m)] + comm
                dur = fwd[s]
                start = max(free[s], dep)
                f_done[(s, m)] = start + dur
            else:
                dep = f_done[(s, m)] if s == S - 1 else b_done[(s + 1, m)] + comm
                dur = bwd[s] + (fwd[s] if (recompute and s < S - 1) else 0.0)
                start = max(free[s], dep)
                b_done[(s, m)] = start + dur
            free[s] = start + dur
            spans.append((s, 0, op, m, start, start + dur))
    makespan = max(free)
    work = [microbatches * (fwd[s] + bwd[s] + (fwd[s] if recompute and s < S - 1 else 0.0)) for s in range(S)]
    return {"makespan_ms": makespan, "bubble_fraction": {str(s): 1.0 - work[s] / makespan for s in range(S)},
            "ideal_ms": max(work), "spans": spans}


def chrome_trace(spans) -> Dict[str, object]:
    ev = [{"name": "%s%d" % (op, m), "ph": "X", "pid": s, "tid": r, "ts": a * 1e3, "dur": (b - a) * 1e3,
           "cat": "forward" if op == OP_F else "backward"} for s, r, op, m, a, b in spans]
    return {"traceEvents": ev, "displayTimeUnit": "ms"}


def cli(argv=None) -> int:
    ap = argparse.ArgumentParser(description="simulate a pipeline schedule")
    ap.add_argument("--fwd", required=True, help="comma-separated forward ms per stage")
    ap.add_argument("--bwd", required=True)
    ap.add_argument("--replicas", default=None)
    ap.add_argument("--comm", type=float, default=0.0)
    ap.add_argument("-n", "--minibatches", type=int, default=64)
    ap.add_argument("--gpipe", type=int, default=0, help="simulate GPipe with this many micro-batches")
    ap.add_argument("--extra-warmup", type=int, default=0,
                    help="additional in-flight minibatches per stage (hides transfer latency)")
    ap.add_argument("--trace", default=None, help="write a Chrome trace JSON here")
    a = ap.parse_args(argv)
    fwd = [float(x) for x in a.fwd.split(",")]
    bwd = [float(x) for x in a.bwd.split(",")]
    if a.gpipe:
        res = simulate_gpipe(fwd, bwd, a.gpipe, a.comm)
    else:
        reps = [int(x) for x in a.replicas.split(",")] if a.replicas else None
        res = simulate_1f1b(fwd, bwd, reps, a.comm, a.minibatches, a.extra_warmup)
    spans = res.pop("spans")
    if a.trace:
        json.dump(chrome_trace(spans), open(a.trace, "w"))
    print(json.dumps(res, indent=1))
    return 0


if __name__ == "__main__":
    raise SystemExit(cli())
