#!/bin/bash
# final single-GPU validation of the round: full GPU test suite, smoke, bench, ncu of the pair GEMM, serial timeline
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu < /dev/null > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/final_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" < /dev/null > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final_smoke.log
timeout 300 python bench.py < /dev/null > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/final_bench_n1.json
timeout 120 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_gemm_pair python tools/ncu_gemm_pair.py < /dev/null > gpurun_out/final_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/final_ncu.log
DDLB_PDL=0 DDLB_WGRAD_OVERLAP=0 timeout 120 python tools/graph_timeline.py serial_final < /dev/null > gpurun_out/final_tl.log 2>&1; echo "timeline rc=$?"; grep "^span" gpurun_out/final_tl.log
