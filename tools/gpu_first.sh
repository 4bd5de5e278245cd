#!/bin/bash
# first-contact run on the B200 box: everything bounded by timeouts, logs into gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,driver_version --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt; free -g | head -2 >> gpurun_out/smi.txt
echo "=== smoke" > gpurun_out/first.log
timeout 300 python __graft_entry__.py smoke >> gpurun_out/first.log 2>&1
echo "exit $?" >> gpurun_out/first.log
echo "=== pytest gpu" >> gpurun_out/first.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x -k "dense" > gpurun_out/pytest_dense.log 2>&1
echo "exit $?" >> gpurun_out/pytest_dense.log
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_gpu_fullsize.py -m gpu -q --timeout 600 -k "not dense" > gpurun_out/pytest_rest.log 2>&1
echo "exit $?" >> gpurun_out/pytest_rest.log
for f in first pytest_dense pytest_rest; do tail -n 5 gpurun_out/$f.log; done
