#!/bin/bash
# GPU visit: tcgen05 phase stamps of one cfg2 step (forward + backward), the -m gpu suite, a bench line
# usage (under gpurun): bash tools/gpu_r2b.sh <tag> [steps: dbg pytest parity benchq bench]
mkdir -p gpurun_out
TAG=$1; shift
STEPS=${@:-dbg pytest benchq}
for step in $STEPS; do
  case $step in
    dbg) timeout 600 python tools/tc2_dbg.py cfg2_mmoe_aliccp_b4096 > gpurun_out/${TAG}_tc2dbg.log 2>&1; tail -3 gpurun_out/${TAG}_tc2dbg.log ;;
    pytest) timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/${TAG}_pytest.log; tail -8 gpurun_out/${TAG}_pytest.log ;;
    parity) timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -40 > gpurun_out/${TAG}_parity.log; tail -8 gpurun_out/${TAG}_parity.log ;;
    benchq) timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cut -c1-400 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err ;;
    bench) timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cut -c1-400 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err ;;
    noside) SWR_SIDE_STREAM=0 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-scopes --no-saturated > gpurun_out/${TAG}_bench_noside.json 2> gpurun_out/${TAG}_bench_noside.err; cut -c1-400 gpurun_out/${TAG}_bench_noside.json ;;
  esac
done
