#!/bin/bash
# Runs the tcgen05 plumbing probe over operand-major / descriptor-variant / shape combinations (one process each,
# so a trapped kernel cannot poison the next case).  Build here: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tc_probe tools/tc_probe.cu
mkdir -p gpurun_out
OUT=gpurun_out/tc_probe.log; : > $OUT
for cfg in "0 0 0 256 96 1 0 0" "0 0 0 256 96 1 0 1" "0 1 0 256 96 1 0 1" "0 0 0 48 128 1 0 1" "0 1 0 64 384 1 0 1" "1 1 0 64 128 1 0 0" "0 0 0 256 384 1 1 0"; do
  timeout 30 tools/bin/tc_probe $cfg >> $OUT 2>&1 || echo "cfg $cfg: exit $?" >> $OUT
done
cat $OUT
