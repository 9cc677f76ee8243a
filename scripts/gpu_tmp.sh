timeout 600 python -m pytest tests/test_gpu_wgrad.py -x -q 2>&1 | tail -5
timeout 600 python scripts/wgrad_check.py 2>&1 | grep -E "custom|not covered|stage 2"
