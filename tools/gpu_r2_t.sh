#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" BENCH_SKIP_E2E=1 timeout 200 python bench.py --steps 40 --warmup 5 --no-torch-baseline < /dev/null > gpurun_out/t_bench_$name.json 2> gpurun_out/t_bench_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/t_bench_%s.json" % n).read().strip().splitlines()[-1])
    print(n, d["value"], "img/s", d["ms_per_step"], "ms; long", d["long_run"]["ms_per_step"])
except Exception as e:
    print(n, "FAILED", e); print(open("gpurun_out/t_bench_%s.err" % n).read()[-800:])
PY
}
run default DDLB_NONE=1
run smallkb8 DDLB_IG_SMALL_MAXKB=8
run smallkb2 DDLB_IG_SMALL_MAXKB=2
run bncluster0 DDLB_BN_CLUSTER_FUSED=0
