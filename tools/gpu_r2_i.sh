#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "avgpool or dropout" < /dev/null > gpurun_out/i_kernels.log 2>&1; echo "kernels rc=$?"; tail -3 gpurun_out/i_kernels.log
timeout 900 python -m pytest tests/test_engine_gpu.py -q -x -s < /dev/null > gpurun_out/i_engine.log 2>&1; echo "engine rc=$?"; tail -15 gpurun_out/i_engine.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" < /dev/null > gpurun_out/i_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/i_smoke.log
