"""Per-node process launcher — the role of ``pipedream-fork/runtime/launch.py:102-173`` (a
``torch.distributed.launch`` clone that passes ``--rank`` / ``--local_rank`` to every child) and
of ``run/run/pipedream_run.sh`` (one background process per local GPU).

    python -m ddlbench_b200.launch --nnodes 1 --node_rank 0 --nproc_per_node 8 \
        [--master_addr 127.0.0.1 --master_port 29500] [--pid-file run.pids] \
        script.py [script args...]

Differences by design: children also get the torchrun environment contract (``RANK``,
``LOCAL_RANK``, ``WORLD_SIZE``, ``MASTER_*``); the launcher records the exact child PIDs in
``--pid-file`` so a run can be stopped without pattern-matching command lines
(``python -m ddlbench_b200.launch --terminate run.pids``; the reference ssh-es ``pkill -f`` to
every machine, ``scripts/terminate_runtime.py:30-31``); when one child fails the rest of the
local group is signalled instead of hanging in a collective.
"""
from __future__ import annotations

import argparse
import os
import signal
import subprocess
import sys
import time
from typing import List


def parse(argv=None):
    ap = argparse.ArgumentParser(description="ddlbench_b200 per-node launcher")
    ap.add_argument("--nnodes", type=int, default=1)
    ap.add_argument("--node_rank", type=int, default=0)
    ap.add_argument("--nproc_per_node", type=int, default=1)
    ap.add_argument("--master_addr", default="127.0.0.1")
    ap.add_argument("--master_port", type=int, default=29500)
    ap.add_argument("--pid-file", default=None)
    ap.add_argument("--no-rank-args", action="store_true",
                    help="do not append --rank/--local_rank (env contract only)")
    ap.add_argument("--terminate", metavar="PIDFILE", default=None)
    ap.add_argument("-m", "--module", action="store_true",
                    help="training_script is a module name (python -m)")
    ap.add_argument("training_script", nargs="?")
    ap.add_argument("training_script_args", nargs=argparse.REMAINDER)
    return ap.parse_args(argv)


def terminate(pid_file: str, grace: float = 5.0) -> int:
    """Signal exactly the PIDs a launcher wrote (TERM, then KILL after ``grace`` seconds)."""
    pids = [int(l) for l in open(pid_file).read().split() if l.strip().isdigit()]
    alive = []
    for p in pids:
        try:
            os.kill(p, signal.SIGTERM)
            alive.append(p)
        except ProcessLookupError:
            pass
    t0 = time.time()
    while alive and time.time() - t0 < grace:
        time.sleep(0.1)
        alive = [p for p in alive if _alive(p)]
    for p in alive:
        try:
            os.kill(p, signal.SIGKILL)
        except ProcessLookupError:
            pass
    print("Done Terminating")
    return len(pids)


def _alive(pid: int) -> bool:
    try:
        os.kill(pid, 0)
        return True
    except ProcessLookupError:
        return False


def main(argv=None) -> int:
    a = parse(argv)
    if a.terminate:
        terminate(a.terminate)
        return 0
    if not a.training_script:
        print("training_script is required", file=sys.stderr)
        return 2
    world = a.nnodes * a.nproc_per_node
    procs: List[subprocess.Popen] = []
    for local in range(a.nproc_per_node):
        rank = a.node_rank * a.nproc_per_node + local
        env = dict(os.environ, MASTER_ADDR=a.master_addr, MASTER_PORT=str(a.master_port),
                   WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(local),
                   LOCAL_WORLD_SIZE=str(a.nproc_per_node))
        cmd = [sys.executable, "-u"] + (["-m"] if a.module else []) + [a.training_script] + \
            list(a.training_script_args)
        if not a.no_rank_args:
            cmd += ["--rank", str(rank), "--local_rank", str(local)]
        procs.append(subprocess.Popen(cmd, env=env))
    if a.pid_file:
        with open(a.pid_file, "w") as f:
            f.write("\n".join(str(p.pid) for p in procs) + "\n")
    rc = 0
    pending = list(procs)
    try:
        while pending:
            for p in list(pending):
                r = p.poll()
                if r is None:
                    continue
                pending.remove(p)
                if r != 0 and rc == 0:
                    rc = r
                    for q in pending:          # exact PIDs of our own children
                        q.terminate()
            time.sleep(0.05)
    except KeyboardInterrupt:
        for q in pending:
            q.terminate()
        rc = 130
    return rc


if __name__ == "__main__":
    raise SystemExit(main())
