#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun flags] -- 'command'   (retries while the pod is busy: rc 3)
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
