#!/bin/bash
# Runs every harness case in its own process (a trapped kernel kills only that case) on the GPU box.
cd "$(dirname "$0")/../footprints_b200/csrc" || exit 1
mkdir -p ../../gpurun_out
H=build/test_harness
N=$($H count)
fail=0
for i in $(seq 0 $((N-1))); do
  timeout 120 $H $i 2>&1 | tail -8
  rc=${PIPESTATUS[0]}
  if [ $rc -ne 0 ]; then echo "case $i rc=$rc"; fail=1; fi
done
if [ "$1" == "perf" ] && [ $fail -eq 0 ]; then
  for i in $(seq 0 10); do timeout 120 $H perf$i 2>&1 | tail -4; done
fi
echo "run_harness fail=$fail"
exit $fail
