timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "merge or host_pipeline or swap or escalation" -p no:cacheprovider --timeout=240 2>&1 | tail -3
bash tools/gpu_profile.sh 2>&1 | tail -5
timeout 120 python tools/gpu_loader_bench.py 1000000 2>&1 | tail -2
