#!/bin/bash
# run.sh-compatible front end: same flags as the reference's run/run/run.sh
# (-b benchmark -f framework -g gpus -n nodes -p loginterval -m model -q queue -s -h),
# but jobs run on the local GPUs instead of being submitted to SLURM.
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
export PYTHONPATH="$ROOT:${PYTHONPATH}"
[ $# -eq 0 ] && exec python -m ddlbench_b200.cli -h
exec python -m ddlbench_b200.cli "$@"
