#!/bin/bash
# One GPU-box visit of round 2.  usage (under gpurun): bash tools/gpu_r2.sh <tag> <steps...>
# steps: buf (buffer-by-buffer parity), pytest (all -m gpu), dbg (tc_debug cfg2), bench, ref, launches, full:<kernel regex>:<count>:<skip>
TAG=$1; shift
mkdir -p gpurun_out
for step in "$@"; do
  case $step in
    buf) timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "every_buffer or golden" 2>&1 | tail -25 > gpurun_out/${TAG}_buf.log; tail -8 gpurun_out/${TAG}_buf.log ;;
    pytest) timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/${TAG}_pytest.log; tail -8 gpurun_out/${TAG}_pytest.log ;;
    dbg) timeout 600 python tests/tc_debug.py cfg2_mmoe_aliccp_b4096 1 > gpurun_out/${TAG}_dbg.log 2>&1; tail -5 gpurun_out/${TAG}_dbg.log ;;
    bench) timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err ;;
    benchq) timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err ;;
    ref) timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_ref.json ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^(fc_|adam|bce|bn_|gather|scatter|head_|pool_|ew_|ln_|mix_|bmv_|select_|sumgrad|colstats)" -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
        python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1 ;;
    full:*) IFS=: read -r _ KRE CNT SKIP <<< "$step"
        timeout 1200 ncu --set full --clock-control none --import-source on -k regex:${KRE} -s ${SKIP:-40} -c ${CNT:-3} -f -o gpurun_out/${TAG}_prof \
        python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1 ;;
    *) echo "unknown step $step" ;;
  esac
done
