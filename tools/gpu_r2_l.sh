#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BENCH=1 OUT=gpurun_out/conv_vs_cudnn_v4.json timeout 400 python tests/igemm_check.py < /dev/null > gpurun_out/l_igemm.log 2>&1; echo "igemm rc=$?"; grep -E "FAIL|SUM|igemm_check|Error|error" gpurun_out/l_igemm.log | head -20
timeout 300 python bench.py --steps 40 --warmup 5 --no-torch-baseline < /dev/null > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/l_bench.json").read().strip().splitlines()[-1])
    print(d["value"], "img/s", d["ms_per_step"], "ms; long", d["long_run"]["ms_per_step"], "e2e", d["e2e"]["value"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/l_bench.err").read()[-2000:])
PY
DDLB_IG_SPLIT=slots timeout 300 python bench.py --steps 40 --warmup 5 --no-torch-baseline < /dev/null > gpurun_out/l_bench_slots.json 2> gpurun_out/l_bench_slots.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/l_bench_slots.json").read().strip().splitlines()[-1])
    print("slots:", d["value"], "img/s", d["ms_per_step"], "ms")
except Exception as e:
    print("bench slots failed", e)
PY
