timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_gpu_configs.py -p no:cacheprovider --timeout=240 > gpurun_out/r02_pytest_small.log 2>&1; tail -4 gpurun_out/r02_pytest_small.log | cut -c1-300
timeout 900 python tools/gpu_sweep9.py 50000000 gpurun_out/r02_sweep9.md > gpurun_out/r02_sweep9.log 2>&1; tail -20 gpurun_out/r02_sweep9.log
