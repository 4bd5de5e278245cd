timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_gpu_configs.py -p no:cacheprovider --timeout=240 > gpurun_out/r02_pytest_small.log 2>&1; tail -6 gpurun_out/r02_pytest_small.log | cut -c1-300
timeout 1500 python -m pytest tests/test_gpu_configs.py -q -s -p no:cacheprovider --timeout=400 > gpurun_out/r02_pytest_configs.log 2>&1; grep -E "passed|failed|Error" gpurun_out/r02_pytest_configs.log | tail -5
bash tools/gpu_perf.sh r02e 2>&1 | tail -30
