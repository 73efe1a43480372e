#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "split_variants or gemm_pair or im2col" < /dev/null > gpurun_out/o_kernels.log 2>&1; echo "kernels rc=$?"; tail -3 gpurun_out/o_kernels.log
timeout 1500 bash tools/sanitize_r2.sh
