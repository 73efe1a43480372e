#!/bin/bash
# Prints the layer summary of every model x dataset on the CPU
# (reference: run/summary/run.sh -> benchmark/network_summary.py).
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
export PYTHONPATH="$ROOT:${PYTHONPATH}"
exec python -m ddlbench_b200.utils.summary "$@"
