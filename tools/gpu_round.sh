#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list, ncu full capture of the top kernels.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [ncu kernel regex]
TAG=${1:-r01}; KRE=${2:-fc_fwd_kernel}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
python bench.py --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:${KRE} -s ${4:-40} -c ${3:-3} -f -o gpurun_out/${TAG}_prof \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
