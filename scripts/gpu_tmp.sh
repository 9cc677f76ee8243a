timeout 300 python scripts/ln_check.py 2>&1 | head -4
timeout 600 python -m pytest tests/test_gpu_layernorm.py tests/test_gpu_model.py -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-240
