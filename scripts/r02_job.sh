#!/bin/bash
# Round-2 GPU job: full GPU suite (kept log), smoke, per-layer times, bench line.  Usage: bash scripts/r02_job.sh <tag>
set -u
export PYTHONUNBUFFERED=1
TAG=${1:-x}
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tee gpurun_out/${TAG}_pytest_gpu.log | grep -vE "^\s*$" | tail -40
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/${TAG}_smoke.log | tail -8
echo "== layer times"
TOP=400 timeout 200 python scripts/layer_times.py > gpurun_out/${TAG}_layer_times.txt 2>&1; head -3 gpurun_out/${TAG}_layer_times.txt
echo "== host time"
timeout 200 python scripts/host_time.py 64 2>&1 | tail -3
timeout 200 python scripts/host_time.py 8 2>&1 | tail -1
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/one_step.py > gpurun_out/${TAG}_one_step.log 2>&1; python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv 2 > gpurun_out/${TAG}_launch_summary.md; head -45 gpurun_out/${TAG}_launch_summary.md
echo "== bench"
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cut -c1-400 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
echo "== bench precise / infer"
timeout 400 python bench.py --precise --no-cpu-baseline > gpurun_out/${TAG}_bench_precise.json 2> gpurun_out/${TAG}_bench_precise.err; cut -c1-260 gpurun_out/${TAG}_bench_precise.json; tail -2 gpurun_out/${TAG}_bench_precise.err
timeout 300 python bench.py --mode infer --no-cpu-baseline > gpurun_out/${TAG}_bench_infer.json 2> gpurun_out/${TAG}_bench_infer.err; cut -c1-260 gpurun_out/${TAG}_bench_infer.json
timeout 300 python bench.py --mode infer --precise --no-cpu-baseline > gpurun_out/${TAG}_bench_infer_precise.json 2> gpurun_out/${TAG}_bench_infer_precise.err; cut -c1-260 gpurun_out/${TAG}_bench_infer_precise.json; tail -2 gpurun_out/${TAG}_bench_infer_precise.err
echo "== bench 512x640 bs16"
timeout 300 python bench.py --hw 512x640 --batch 16 --no-cpu-baseline --no-precise > gpurun_out/${TAG}_bench_512x640_bs16.json 2> gpurun_out/${TAG}_bench_512.err; cut -c1-260 gpurun_out/${TAG}_bench_512x640_bs16.json; tail -2 gpurun_out/${TAG}_bench_512.err
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 2 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_ref.err; cut -c1-400 gpurun_out/${TAG}_bench_reference_arm.json; tail -2 gpurun_out/${TAG}_ref.err
