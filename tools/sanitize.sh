#!/bin/bash
# Run the kernel numerics tests under compute-sanitizer (memcheck, racecheck, synccheck, initcheck).
# The reference uses no sanitizer anywhere.  One GPU; run it through gpurun:
#   gpurun --timeout 1500 -- 'tools/sanitize.sh memcheck racecheck'
# Results go to gpurun_out/sanitize_<tool>.log; a non-zero exit means a tool reported errors.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOLS=("$@"); [ ${#TOOLS[@]} -eq 0 ] && TOOLS=(memcheck racecheck synccheck)
# a focused selection keeps the (slow, serialised) sanitizer run within minutes
SEL='test_sgd_flat or test_xent or test_bn_forward_backward or test_normalize_u8 or test_revert_varlen_and_mask or test_global_avg_pool or test_maxpool or test_l2norm'
rc=0
for t in "${TOOLS[@]}"; do
  echo "== compute-sanitizer --tool $t"
  compute-sanitizer --tool "$t" --error-exitcode 9 --launch-timeout 120 \
    python -m pytest tests/test_kernels_gpu.py -x -q -k "$SEL" > "gpurun_out/sanitize_$t.log" 2>&1
  r=$?
  tail -4 "gpurun_out/sanitize_$t.log"
  grep -c "ERROR SUMMARY: 0 errors" "gpurun_out/sanitize_$t.log" > /dev/null || true
  [ $r -ne 0 ] && rc=$r
done
exit $rc
