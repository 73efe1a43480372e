#!/bin/bash
# round-2 bring-up B: igemm v2 (2 CTAs/SM, chunked) numerics + graph-timed per-layer table, A/B bench, launch list
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BENCH=1 OUT=gpurun_out/conv_vs_cudnn.json timeout 600 python tests/igemm_check.py < /dev/null > gpurun_out/igemm_check.log 2>&1
echo "igemm_check rc=$?"; tail -3 gpurun_out/igemm_check.log
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
run ig1_nooverlap DDLB_IGEMM=1 DDLB_WGRAD_OVERLAP=0
run ig1_3blk DDLB_IGEMM=1 DDLB_BN_BWD_APPLY_BLOCKS=3
run ig0 DDLB_IGEMM=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/launches_ig1.csv python tools/step_profile.py < /dev/null > gpurun_out/ncu_ig1.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_ig1.csv 40 | tee gpurun_out/launches_ig1.txt | head -45
