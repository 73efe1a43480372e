#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -x -q < /dev/null > gpurun_out/pytest_multigpu_n4.log 2>&1; echo "pytest multigpu rc=$?"; tail -5 gpurun_out/pytest_multigpu_n4.log
pd() {
  name=$1; cfg=$2; shift 2
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus 4 --config $cfg --steps 24 --warmup 3 < /dev/null > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/bench_%s.json" % n).read().strip().splitlines()[-1])
    print(n, d.get("value"), "img/s", d.get("ms_per_step"), "ms", d.get("config", {}).get("epoch_throughputs"), d.get("unavailable"))
except Exception as e:
    print(n, "FAILED", e); print(open("gpurun_out/bench_%s.err" % n).read()[-1800:])
PY
}
pd pd_r50_4_eager pd_r50_4 DDLB_PIPE_GRAPHS=0
pd pd_r50_4_graphs pd_r50_4 DDLB_PIPE_GRAPHS=1
pd pd_r50_4_graphs_symm pd_r50_4 DDLB_PIPE_GRAPHS=1 DDLB_P2P=symm
pd gpipe_vgg16 gpipe_vgg16_c10_4 DDLB_NONE=1
