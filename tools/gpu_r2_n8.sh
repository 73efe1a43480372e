#!/bin/bash
# One 8-GPU call: DP bench at N=8, multi-GPU pytest (collectives, fused AR+SGD, K4 stashing replicas),
# the PipeDream bench configs and the all-reduce sweep.  Every step has its own timeout.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tr() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$2" --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "${@:3}"; }
tr 240 8 bench.py --gpus 8 --steps 20 --warmup 5 --no-torch-baseline < /dev/null > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?"
tr 200 8 bench.py --gpus 8 --config pd_r50_4x2 --steps 24 < /dev/null > gpurun_out/bench_pd_r50_4x2.json 2> gpurun_out/bench_pd_r50_4x2.err; echo "pd 4x2 rc=$?"
tr 200 8 bench.py --gpus 8 --config pd_r152_8 --steps 24 < /dev/null > gpurun_out/bench_pd_r152_8.json 2> gpurun_out/bench_pd_r152_8.err; echo "pd r152 rc=$?"
CUDA_VISIBLE_DEVICES=0,1,2,3 tr 200 4 bench.py --gpus 4 --config pd_r50_4 --steps 24 < /dev/null > gpurun_out/bench_pd_r50_4.json 2> gpurun_out/bench_pd_r50_4.err; echo "pd r50_4 rc=$?"
python - <<'PY'
import json
for n in ("n8", "pd_r50_4x2", "pd_r152_8", "pd_r50_4"):
    try:
        d = json.loads(open("gpurun_out/bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, d.get("value"), d.get("ms_per_step"), d.get("e2e", {}).get("value"), d.get("checks"), d.get("config", {}).get("epoch_throughputs"), d.get("unavailable"))
    except Exception as e:
        print(n, "FAILED", e); print(open("gpurun_out/bench_%s.err" % n).read()[-1200:])
PY
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 400 python -m pytest tests/test_multigpu.py -q < /dev/null > gpurun_out/pytest_multigpu.log 2>&1; echo "pytest multigpu rc=$?"; tail -4 gpurun_out/pytest_multigpu.log
SWEEP=1 tr 200 8 tests/dist_check_gpu.py < /dev/null > gpurun_out/dist8.log 2>&1; echo "dist8 rc=$?"; grep -E "OK|FAIL|bus_GBps" gpurun_out/dist8.log | tail -16
BLOCKS=48 tr 150 8 bench/allreduce_bandwidth.py < /dev/null > gpurun_out/ar_bw_n8.log 2>&1; echo "ar_bw rc=$?"; grep "^{" gpurun_out/ar_bw_n8.log | tail -8
