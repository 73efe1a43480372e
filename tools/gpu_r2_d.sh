#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
QUICK=0 timeout 300 python tests/igemm_check.py < /dev/null > gpurun_out/igemm_check.log 2>&1; echo "igemm_check rc=$?"; tail -1 gpurun_out/igemm_check.log
DDLB_IGEMM=1 timeout 300 python tools/graph_timeline.py ig1 2>&1 | grep -v Warning | tail -30
DDLB_IGEMM=0 timeout 300 python tools/graph_timeline.py ig0 2>&1 | grep -v Warning | tail -30
