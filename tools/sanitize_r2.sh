#!/bin/bash
# compute-sanitizer over the round-2 kernel paths (tests/sanitize_targets.py).  One GPU:
#   gpurun --timeout 1200 -- 'tools/sanitize_r2.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rc=0
for t in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $t"
  timeout 420 compute-sanitizer --tool "$t" --error-exitcode 9 --launch-timeout 120 \
    python tests/sanitize_targets.py > "gpurun_out/sanitize_r2_$t.log" 2>&1
  r=$?
  echo "rc=$r"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize targets done|Error|error" "gpurun_out/sanitize_r2_$t.log" | head -8
  [ $r -ne 0 ] && rc=$r
done
# the distributed-shared-memory flavour of the cluster split-K exchange
for t in racecheck; do
  echo "== compute-sanitizer --tool $t (DDLB_IG_SPLIT=dsmem)"
  DDLB_IG_SPLIT=dsmem timeout 420 compute-sanitizer --tool "$t" --error-exitcode 9 --launch-timeout 120 \
    python tests/sanitize_targets.py > "gpurun_out/sanitize_r2_dsmem_$t.log" 2>&1
  r=$?
  echo "rc=$r"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize targets done|Error|error" "gpurun_out/sanitize_r2_dsmem_$t.log" | head -8
  [ $r -ne 0 ] && rc=$r
done
exit $rc
