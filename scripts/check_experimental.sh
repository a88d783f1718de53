#!/bin/bash
# First GPU job of the next round: validate and time the opt-in kernels (run from the repo root on a B200 box):
#   gpurun --timeout 900 -- 'bash scripts/check_experimental.sh > gpurun_out/experimental.log 2>&1; tail -40 gpurun_out/experimental.log'
# 1. parity of the row-streaming kernel and of the env-selected dispatch variants (opt-in tests)
# 2. per-layer times of one training step with the default dispatch and with FP_ROWSTREAM=1
set -u
export PYTHONUNBUFFERED=1
echo "== opt-in parity tests"
FP_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -q -m gpu -x 2>&1 | tail -15
echo "== layer times, default dispatch"
TOP=400 timeout 200 python scripts/layer_times.py 2>&1 | grep -E "step ms|^\{|fwd|dgrad" | head -24
echo "== layer times, FP_ROWSTREAM=1"
FP_ROWSTREAM=1 TOP=400 timeout 200 python scripts/layer_times.py 2>&1 | grep -E "step ms|^\{|fwd|dgrad" | head -24
echo "== layer times, FP_WGRAD_ROLL=1"
FP_WGRAD_ROLL=1 TOP=400 timeout 200 python scripts/layer_times.py 2>&1 | grep -E "step ms|^\{|wgrad" | head -24
echo "== end-to-end parity with FP_ROWSTREAM=1 FP_WGRAD_ROLL=1"
FP_ROWSTREAM=1 FP_WGRAD_ROLL=1 timeout 400 python -m pytest tests/test_gpu_forward.py tests/test_gpu_train.py -q -m gpu -x 2>&1 | tail -5
