#!/bin/bash
# Round-2 first GPU job: kept pytest log, smoke, experimental kernels, bench lines.
set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tee gpurun_out/r02_pytest_gpu.log | tail -15
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/r02_smoke.log | tail -8
echo "== experimental"
FP_TEST_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gpu_experimental.py -q -m gpu 2>&1 | tee gpurun_out/r02_experimental.log | tail -40
echo "== layer times default"
TOP=400 timeout 200 python scripts/layer_times.py > gpurun_out/r02_layer_times_default.txt 2>&1; tail -3 gpurun_out/r02_layer_times_default.txt
echo "== layer times rowstream"
FP_ROWSTREAM=1 TOP=400 timeout 200 python scripts/layer_times.py > gpurun_out/r02_layer_times_rowstream.txt 2>&1; tail -3 gpurun_out/r02_layer_times_rowstream.txt
echo "== layer times wgrad roll"
FP_WGRAD_ROLL=1 TOP=400 timeout 200 python scripts/layer_times.py > gpurun_out/r02_layer_times_roll.txt 2>&1; tail -3 gpurun_out/r02_layer_times_roll.txt
echo "== bench"
timeout 400 python bench.py > gpurun_out/r02_bench0.json 2> gpurun_out/r02_bench0.err; cat gpurun_out/r02_bench0.json | cut -c1-600
