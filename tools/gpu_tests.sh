#!/bin/bash
# GPU box: the -m gpu suite in two parts (small cases first, then the BASELINE-size configs), logs under gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_gpu_configs.py -p no:cacheprovider --timeout=240 > gpurun_out/pytest_small.log 2>&1
echo "small rc=$?" >> gpurun_out/pytest_small.log
tail -30 gpurun_out/pytest_small.log
timeout 2400 python -m pytest tests/test_gpu_configs.py -q -s --maxfail=6 -p no:cacheprovider --timeout=240 > gpurun_out/pytest_configs.log 2>&1
echo "configs rc=$?" >> gpurun_out/pytest_configs.log
grep -E "parity\]|passed|failed|Error|error" gpurun_out/pytest_configs.log | tail -40
