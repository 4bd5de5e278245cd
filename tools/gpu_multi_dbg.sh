#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
SWAT_DEBUG=1 NCCL_DEBUG=WARN timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/dist_check_$N.log 2>&1
echo "dist check exit $?"; grep -E "rank|swat dist|world=|rror" gpurun_out/dist_check_$N.log | tail -40
