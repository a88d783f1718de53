#!/bin/bash
# ncu --set full of the dominant kernels of one training step (NVTX-filtered), one GPU.  Usage: bash scripts/r02_ncu.sh <tag>
# The .ncu-rep stays on the box when it is too large to travel; the raw / details pages are exported there as CSV.
set -u
TAG=${1:-x}
export PYTHONUNBUFFERED=1 FP_NVTX=1 FP_OVERLAP_WGRAD=0
mkdir -p gpurun_out
INC=""
for r in fwd:mask.o4.c1 fwd:mask.o4.c2 dgrad:mask.o4.c1 dgrad:mask.o4.c2 wgrad:mask.o4.c1 wgrad:mask.o4.c2 fwd:mask.block4.post1 fwd:enc1.1.conv1 fwd:enc3.1.conv1 wgrad:enc3.1.conv1 fwd:mask.head4; do
  INC="$INC --nvtx-include $r/"
done
timeout 1500 ncu --set full --clock-control none --import-source on --nvtx $INC -k regex:"conv_|head_wgrad" -c 12 -o /tmp/${TAG}_prof -f python scripts/one_step.py > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ncu -i /tmp/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_prof.ncu-rep --page details --csv > gpurun_out/${TAG}_ncu_details.csv 2>/dev/null
ls -la /tmp/${TAG}_prof.ncu-rep gpurun_out/${TAG}_ncu_raw.csv gpurun_out/${TAG}_ncu_details.csv
sz=$(stat -c %s /tmp/${TAG}_prof.ncu-rep); if [ "$sz" -lt 40000000 ]; then cp /tmp/${TAG}_prof.ncu-rep gpurun_out/; fi
