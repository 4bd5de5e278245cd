timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_gpu_configs.py -p no:cacheprovider --timeout=240 > gpurun_out/r02_pytest_small.log 2>&1; tail -15 gpurun_out/r02_pytest_small.log | cut -c1-300
timeout 120 python tools/gpu_loader_bench.py 1000000 2>&1 | tail -3
