timeout 300 python tools/gpu_lock_exp.py 10000000 2>&1 | tail -22
timeout 600 python -m pytest tests/test_gpu_property.py tests/test_gpu_parity.py -q -k "property or random_draws or synonym or grouped or partitioned_escalation" -p no:cacheprovider --timeout=300 2>&1 | tail -5
