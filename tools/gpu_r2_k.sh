#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "avgpool or dropout or im2col" < /dev/null > gpurun_out/k_kernels.log 2>&1; echo "kernels rc=$?"; tail -3 gpurun_out/k_kernels.log
DDLB_PDL=0 DDLB_WGRAD_OVERLAP=0 timeout 300 python tools/graph_timeline.py serial < /dev/null > gpurun_out/k_tl_serial.log 2>&1; echo "timeline rc=$?"; head -3 gpurun_out/k_tl_serial.log
timeout 300 python bench.py --steps 40 --warmup 5 --no-torch-baseline < /dev/null > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/k_bench.json").read().strip().splitlines()[-1])
print(d["value"], "img/s", d["ms_per_step"], "ms; long", d["long_run"]["ms_per_step"], "e2e", d["e2e"]["value"])
PY
