#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for ov in 1 0; do
DDLB_SGD_OVERLAP=$ov timeout 300 python bench.py --steps 40 --warmup 5 --no-torch-baseline < /dev/null > gpurun_out/s_bench_ov$ov.json 2> gpurun_out/s_bench_ov$ov.err; python - $ov <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/s_bench_ov%s.json" % f).read().strip().splitlines()[-1])
    print("sgd_overlap=%s:" % f, d["value"], "img/s", d["ms_per_step"], "ms; long", d["long_run"]["ms_per_step"], "launches", d.get("gpu_launches"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/s_bench_ov%s.err" % f).read()[-2000:])
PY
done
