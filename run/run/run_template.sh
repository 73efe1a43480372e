#!/bin/bash
# Job body with the reference's 11 positional arguments (run/run/run_template.sh:14-28):
#   $1 benchmark  $2 framework  $3 GPUs per node  $4 nodes  $5 CPUs per GPU  $6 model
#   $7 queue time (ignored: no batch system)  $8 output file  $9 queue name (ignored)
#   $10 print interval  $11 use synthetic (0 = synthetic, 1 = real data from $DATADIR)
# "all" for benchmark / framework / model expands to the same lists as the reference.  The job
# runs on the local node; the environment contract (CORES_GPU, LOGINTER, EPOCHS, DATADIR, ...)
# is exported for the worker processes exactly as the sbatch body does.
set -u
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
export PYTHONPATH="$ROOT:${PYTHONPATH:-}"
if [ $# -lt 11 ]; then
  sed -n 2,9p "$0"; exit 1
fi
BENCH=$1; FRAMEWORK=$2; GPUS=$3; NODES=$4; export CORES_GPU=$5; MODEL=$6; OUTFILE=$8
export LOGINTER=${10}; USE_SYNTH=${11}
export EPOCHS=${EPOCHS:-3}
ARGS=(-b "$BENCH" -f "$FRAMEWORK" -g "$GPUS" -n "$NODES" -p "$LOGINTER" -m "$MODEL" --epochs "$EPOCHS")
[ "$USE_SYNTH" = "1" ] && ARGS+=(-s)
[ -n "${STEPS:-}" ] && ARGS+=(--steps "$STEPS")      # optional cap on iterations per epoch
mkdir -p "$(dirname "$OUTFILE")"
python -m ddlbench_b200.cli "${ARGS[@]}" --out "$(dirname "$OUTFILE")" 2>&1 | tee "$OUTFILE"
