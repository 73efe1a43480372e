#!/bin/bash
# final 4-GPU check: multi-GPU tests (K4 fused replica reduction, pipelines) + the pipeline benches
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -x -q < /dev/null > gpurun_out/pytest_multigpu_n4b.log 2>&1; echo "pytest multigpu rc=$?"; tail -5 gpurun_out/pytest_multigpu_n4b.log
pd() {
  name=$1; cfg=$2; shift 2
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus 4 --config $cfg --steps 24 --warmup 3 < /dev/null > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/bench_%s.json" % n).read().strip().splitlines()[-1])
    print(n, d.get("value"), "img/s", d.get("ms_per_step"), "ms", d.get("unavailable"))
except Exception as e:
    print(n, "FAILED", e); print(open("gpurun_out/bench_%s.err" % n).read()[-1800:])
PY
}
pd pd_r50_4_final pd_r50_4 DDLB_NONE=1
pd gpipe_vgg16_final gpipe_vgg16_c10_4 DDLB_NONE=1
pd dp4_final dp DDLB_NONE=1
