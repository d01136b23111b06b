#!/bin/bash
# 2-GPU visit: peer-memory shard checks (dense and row-lazy Adam on the shards) and the sharded bench lines
mkdir -p gpurun_out
TAG=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29541 tests/dist_p2p_check.py > gpurun_out/${TAG}_p2p_dense.log 2>&1; grep -o "DIST_P2P_CHECK_OK.*" gpurun_out/${TAG}_p2p_dense.log | cut -c1-300 || tail -5 gpurun_out/${TAG}_p2p_dense.log
SWR_LAZY_SHARD_MIN=0 timeout 200 $TR --master-port 29542 tests/dist_p2p_check.py > gpurun_out/${TAG}_p2p_lazy.log 2>&1; grep -o "DIST_P2P_CHECK_OK.*" gpurun_out/${TAG}_p2p_lazy.log | cut -c1-300 || tail -15 gpurun_out/${TAG}_p2p_lazy.log
timeout 300 $TR --master-port 29543 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_cfg2_n2.json 2> gpurun_out/${TAG}_cfg2_n2.err; cut -c1-300 gpurun_out/${TAG}_cfg2_n2.json
timeout 400 $TR --master-port 29544 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline --workload cfg4b_ppnet_aliccp85m_b4096 > gpurun_out/${TAG}_cfg4b_n2.json 2> gpurun_out/${TAG}_cfg4b_n2.err; cut -c1-300 gpurun_out/${TAG}_cfg4b_n2.json; tail -2 gpurun_out/${TAG}_cfg4b_n2.err | cut -c1-300
