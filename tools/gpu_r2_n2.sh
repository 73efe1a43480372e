#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check_gpu.py < /dev/null > gpurun_out/dist2.log 2>&1; echo "dist_check rc=$?"
grep -E "OK|FAIL|bus_GBps" gpurun_out/dist2.log | tail -24
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-torch-baseline < /dev/null > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
    print("n2", d["value"], d["ms_per_step"], "long", d["long_run"]["ms_per_step"], "e2e", d["e2e"].get("value"), "checks", d.get("checks"))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/bench_n2.err").read()[-2500:])
PY
