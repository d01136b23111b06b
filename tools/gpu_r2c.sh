#!/bin/bash
mkdir -p gpurun_out
TAG=$1; shift
timeout 200 python tools/stress_forward.py 2>&1 | tail -2 | cut -c1-300 > gpurun_out/${TAG}_stress.log; cat gpurun_out/${TAG}_stress.log
bash tools/gpu_r2b.sh $TAG ${@:-dbg parity benchq}
