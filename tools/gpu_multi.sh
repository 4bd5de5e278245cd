#!/bin/bash
# multi-GPU: distributed parity check + sharded bench (N = $1 GPUs)
N=${1:-2}
mkdir -p gpurun_out
for ROWS in 4000000 16000000; do
CHECK_ROWS=$ROWS timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/dist_check_${N}_$ROWS.log 2>&1
echo "dist check rows=$ROWS exit $?"; grep -E "world=|rror" gpurun_out/dist_check_${N}_$ROWS.log | tail -6
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err
echo "bench exit $?"; tail -c 2200 gpurun_out/bench_g$N.json; tail -n 4 gpurun_out/bench_g$N.err | cut -c1-300
