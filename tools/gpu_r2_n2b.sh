#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-torch-baseline < /dev/null > gpurun_out/bench_n2b.json 2> gpurun_out/bench_n2b.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_n2b.json").read().strip().splitlines()[-1])
    print("n2", d["value"], d["ms_per_step"], "long", d["long_run"]["ms_per_step"], "e2e", d["e2e"].get("value"), "checks", d.get("checks"))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/bench_n2b.err").read()[-2500:])
PY
timeout 600 python -m pytest tests/test_multigpu.py -q -x < /dev/null > gpurun_out/pytest_multigpu_n2b.log 2>&1; echo "multigpu rc=$?"; tail -5 gpurun_out/pytest_multigpu_n2b.log
