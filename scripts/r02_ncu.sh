#!/bin/bash
# ncu --set full of the dominant kernels of one training step (NVTX-filtered), one GPU.  Usage: bash scripts/r02_ncu.sh <tag>
set -u
TAG=${1:-x}
export PYTHONUNBUFFERED=1 FP_NVTX=1
mkdir -p gpurun_out
INC=""
for r in fwd:mask.o4.c1 fwd:mask.o4.c2 dgrad:mask.o4.c1 dgrad:mask.o4.c2 wgrad:mask.o4.c1 wgrad:mask.o4.c2 fwd:mask.block4.post1 fwd:mask.block4.post2 dgrad:mask.block4.post1 wgrad:mask.block4.post1 fwd:enc1.1.conv1 fwd:enc2.1.conv1 fwd:enc3.1.conv1 fwd:enc4.1.conv1 wgrad:enc3.1.conv1 dgrad:enc3.1.conv1 wgrad:mask.head4 fwd:mask.head4; do
  INC="$INC --nvtx-include $r/"
done
timeout 1500 ncu --set full --clock-control none --import-source on --nvtx $INC -k regex:"conv_|wgrad_reduce" -o gpurun_out/${TAG}_prof -f python scripts/one_step.py > gpurun_out/${TAG}_ncu.log 2>&1
tail -5 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out/${TAG}_prof.ncu-rep
