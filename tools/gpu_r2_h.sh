#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 40 --warmup 5 --no-torch-baseline < /dev/null > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/bench_%s.json" % n).read().strip().splitlines()[-1])
    print(n, d["value"], "img/s", d["ms_per_step"], "ms long", d["long_run"]["ms_per_step"], "e2e", d["e2e"]["value"], "thr", d["e2e"].get("decode_threads"))
except Exception as e:
    print(n, "FAILED", e); print(open("gpurun_out/bench_%s.err" % n).read()[-1500:])
PY
}
run prio1 DDLB_MAIN_PRIO=1 BENCH_DECODE_THREADS=64
run prio0 DDLB_MAIN_PRIO=0 BENCH_DECODE_THREADS=96
run prio1_nooverlap DDLB_MAIN_PRIO=1 DDLB_WGRAD_OVERLAP=0 BENCH_DECODE_THREADS=48
