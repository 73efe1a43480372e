#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for mode in cluster off; do
DDLB_IG_SPLIT=$mode timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_smallm_$mode python tools/ncu_igemm.py l4.b1.conv1 l3.b1.conv1 l3.b1.conv2 < /dev/null > gpurun_out/m_ncu_$mode.log 2>&1; echo "ncu $mode rc=$?"; tail -2 gpurun_out/m_ncu_$mode.log
done
ls -la gpurun_out/*.ncu-rep
