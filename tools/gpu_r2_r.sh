#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "bn_apply_finalizes or test_bn" < /dev/null > gpurun_out/r_kernels.log 2>&1; echo "kernels rc=$?"; tail -5 gpurun_out/r_kernels.log
timeout 900 python -m pytest tests/test_engine_gpu.py -q -x < /dev/null > gpurun_out/r_engine.log 2>&1; echo "engine rc=$?"; tail -5 gpurun_out/r_engine.log
for fin in 1 0; do
DDLB_BN_APPLY_FIN=$fin timeout 300 python bench.py --steps 40 --warmup 5 --no-torch-baseline < /dev/null > gpurun_out/r_bench_fin$fin.json 2> gpurun_out/r_bench_fin$fin.err; python - $fin <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/r_bench_fin%s.json" % f).read().strip().splitlines()[-1])
    print("apply_fin=%s:" % f, d["value"], "img/s", d["ms_per_step"], "ms; long", d["long_run"]["ms_per_step"], "launches", d.get("gpu_launches"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r_bench_fin%s.err" % f).read()[-2000:])
PY
done
