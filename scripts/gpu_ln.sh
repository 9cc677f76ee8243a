#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layernorm.py -x -q > gpurun_out/ln_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ln_pytest.log
tail -n 12 gpurun_out/ln_pytest.log
timeout 300 python scripts/ln_check.py > gpurun_out/ln_check.log 2>&1; cat gpurun_out/ln_check.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/ln_pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ln_pytest_all.log
tail -n 8 gpurun_out/ln_pytest_all.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ln_bench.log 2>&1; cut -c1-400 gpurun_out/ln_bench.log
