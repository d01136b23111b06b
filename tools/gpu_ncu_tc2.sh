#!/bin/bash
# ncu --set full (source-level stall sampling) of the tcgen05 launches of one cfg2 training step
mkdir -p gpurun_out
TAG=$1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fc_tc2 -s 22 -c 11 -f -o gpurun_out/${TAG}_tc2 \
  python tools/tc2_step.py cfg2_mmoe_aliccp_b4096 > gpurun_out/${TAG}_tc2_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_tc2_ncu.log
