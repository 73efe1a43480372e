#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python tools/avgdbg.py < /dev/null > gpurun_out/j_avgdbg.log 2>&1; echo "avgdbg rc=$?"; tail -12 gpurun_out/j_avgdbg.log
BENCH=1 OUT=gpurun_out/gemm_tcgen05_vs_cublas.json timeout 300 python tests/gemm_pair_check.py < /dev/null > gpurun_out/j_pair.log 2>&1; echo "pair rc=$?"; tail -40 gpurun_out/j_pair.log
