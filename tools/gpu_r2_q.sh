#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/wgrad_trace.py l4.b1.conv2 l3.b1.conv1 2>&1 | tee gpurun_out/wgrad_trace_v4.log | grep -E "wgrad:|acc ready  |epilogue|after bar"
BENCH=1 OUT=gpurun_out/conv_vs_cudnn_v8.json timeout 400 python tests/igemm_check.py < /dev/null > gpurun_out/q_igemm.log 2>&1; echo "igemm rc=$?"; grep -E "FAIL|SUM|igemm_check|Error|error" gpurun_out/q_igemm.log | head -20
timeout 300 python bench.py --steps 40 --warmup 5 --no-torch-baseline < /dev/null > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/q_bench.json").read().strip().splitlines()[-1])
    print(d["value"], "img/s", d["ms_per_step"], "ms; long", d["long_run"]["ms_per_step"], "e2e", d["e2e"]["value"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/q_bench.err").read()[-2000:])
PY
