#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
CHECK_ROWS=16777216 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/dist_check_$N.log 2>&1
echo "dist check exit $?"; grep -E "world=|rror" gpurun_out/dist_check_$N.log | tail -4
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err
echo "bench exit $?"; tail -c 1500 gpurun_out/bench_g$N.json
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --rows 50000000 --classes 1000 --t2t-only --no-e2e > gpurun_out/bench_cfg4_g$N.json 2> gpurun_out/bench_cfg4_g$N.err
echo "cfg4 exit $?"; tail -c 1500 gpurun_out/bench_cfg4_g$N.json; tail -n 3 gpurun_out/bench_cfg4_g$N.err | cut -c1-300
