#!/bin/bash
# round-2 bring-up A: numerics + per-layer timing of the igemm family, then A/B bench
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BENCH=1 OUT=gpurun_out/conv_vs_cudnn.json timeout 600 python tests/igemm_check.py < /dev/null > gpurun_out/igemm_check.log 2>&1
echo "igemm_check rc=$?" | tee -a gpurun_out/igemm_check.log
tail -5 gpurun_out/igemm_check.log
for cfg in "DDLB_IGEMM=1" "DDLB_IGEMM=0" "DDLB_IGEMM=1 DDLB_BN_BWD_APPLY_BLOCKS=3"; do
  name=$(echo $cfg | tr ' =' '__')
  env $cfg timeout 300 python bench.py --steps 40 --warmup 5 < /dev/null > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  echo "$cfg rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$name.json").read().strip().splitlines()[-1])
    print("$cfg", d["value"], "img/s", d["ms_per_step"], "ms loss", d["e2e"]["last_loss"], "launches", d["launches_per_step_in_graph"])
except Exception as e:
    print("$cfg FAILED", e); print(open("gpurun_out/bench_$name.err").read()[-1500:])
PY
done
