#!/bin/bash
# 2-GPU data-parallel check: bench.py at N=1 and under torchrun at N=2 (the launch line the driver uses)
set -u
export PYTHONUNBUFFERED=1
TAG=${1:-x}
N=${2:-2}
mkdir -p gpurun_out
timeout 300 python bench.py --gpus 1 --no-cpu-baseline --no-precise > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err; cut -c1-170 gpurun_out/${TAG}_n1.json
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline --no-precise > gpurun_out/${TAG}_n${N}.json 2> gpurun_out/${TAG}_n${N}.err
grep -h '"metric"' gpurun_out/${TAG}_n${N}.json | cut -c1-170; tail -2 gpurun_out/${TAG}_n${N}.err
