#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3 = nothing charged).  Usage: gpurun_retry.sh <tag> <timeout> <cmd>
TAG=$1; TMO=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$TMO" -- "$@" > gpurun_out/${TAG}_call.txt 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" gpurun_out/${TAG}_call.txt; then exit $rc; fi
  sleep 120
done
exit 3
