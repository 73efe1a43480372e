#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BENCH=1 OUT=gpurun_out/conv_vs_cudnn.json timeout 600 python tests/igemm_check.py < /dev/null > gpurun_out/igemm_check.log 2>&1
echo "igemm_check rc=$?"; grep "FAIL" gpurun_out/igemm_check.log | head; tail -2 gpurun_out/igemm_check.log
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 40 --warmup 5 < /dev/null > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/bench_%s.json" % n).read().strip().splitlines()[-1])
    print(n, d["value"], "img/s", d["ms_per_step"], "ms loss", d["e2e"]["last_loss"], "launches", d["launches_per_step_in_graph"])
except Exception as e:
    print(n, "FAILED", e); print(open("gpurun_out/bench_%s.err" % n).read()[-1500:])
PY
}
run ig1 DDLB_IGEMM=1
run ig1_small9 DDLB_IGEMM=1 DDLB_IG_SMALL_MAXKB=9
run ig1_3blk DDLB_IGEMM=1 DDLB_BN_BWD_APPLY_BLOCKS=3
run ig1_cluster DDLB_IGEMM=1 DDLB_BN_CLUSTER=1
DDLB_IGEMM=1 timeout 300 python tools/graph_timeline.py ig1 2>&1 | grep -v -i warn | tail -14
