#!/bin/bash
# Per-node PipeDream launcher with the reference's getopts flags (run/run/pipedream_run.sh:10-58):
#  -m master ip   -n model name   -t total GPUs   -l GPUs per node   -c parallel configuration
#  (mp_conf / dp_conf / hybrid_conf .json path)   -b backend (ignored: nccl on GPUs, gloo on CPU)
#  -w CPU workers per GPU   -s use synthetic data (0/1)   -i node index   -o log interval
#  -a batch size   -d data path   -e epochs   -r highres (0/1)
# One worker per local GPU is started through ddlbench_b200.launch (which records the exact
# child PIDs so the run can be terminated without pattern-matching process names).
set -u
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
export PYTHONPATH="$ROOT:${PYTHONPATH:-}"
MASTER=127.0.0.1; MODEL=resnet50; TOTAL=1; LOCAL=1; CONF=""; WORKERS=4; SYNTH=1; NODE=0
LOGI=25; BATCH=""; DATA="${DATADIR:-}"; EP=3; HIGHRES=0; BENCH=${BENCHMARK:-imagenet}
while getopts "m:n:t:l:c:b:w:s:i:o:a:d:e:r:h" arg; do
  case $arg in
    m) MASTER=$OPTARG;; n) MODEL=$OPTARG;; t) TOTAL=$OPTARG;; l) LOCAL=$OPTARG;; c) CONF=$OPTARG;;
    b) ;; w) WORKERS=$OPTARG;; s) SYNTH=$OPTARG;; i) NODE=$OPTARG;; o) LOGI=$OPTARG;;
    a) BATCH=$OPTARG;; d) DATA=$OPTARG;; e) EP=$OPTARG;; r) HIGHRES=$OPTARG;;
    h|*) sed -n 2,9p "$0"; exit 0;;
  esac
done
[ "$HIGHRES" = "1" ] && BENCH=highres
export CORES_GPU=$WORKERS DATADIR=$DATA LOGINTER=$LOGI EPOCHS=$EP
ARGS=(-b "$BENCH" -f pipedream -g "$LOCAL" -n $(( TOTAL / LOCAL )) -p "$LOGI" -m "$MODEL" --epochs "$EP" --worker)
[ -n "$CONF" ] && ARGS+=(--config-path "$CONF")
[ -n "$BATCH" ] && ARGS+=(--batch-size "$BATCH")
[ "$SYNTH" = "0" ] && ARGS+=(-s)
[ -n "${STEPS:-}" ] && ARGS+=(--steps "$STEPS")      # optional cap on iterations per epoch
exec python -m ddlbench_b200.launch --nnodes $(( TOTAL / LOCAL )) --node_rank "$NODE" \
  --nproc_per_node "$LOCAL" --master_addr "$MASTER" --no-rank-args --pid-file "pipedream_node$NODE.pids" \
  -m ddlbench_b200.cli "${ARGS[@]}"
