#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nproc; free -g | head -2
BENCH_NVJPEG=1 timeout 600 python bench.py --steps 20 --warmup 5 < /dev/null > gpurun_out/bench_new_n1.json 2> gpurun_out/bench_new_n1.err; echo "ours rc=$?"
tail -c 3000 gpurun_out/bench_new_n1.json; tail -5 gpurun_out/bench_new_n1.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 < /dev/null > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "ref rc=$?"
tail -c 2500 gpurun_out/bench_ref_n1.json; tail -5 gpurun_out/bench_ref_n1.err
