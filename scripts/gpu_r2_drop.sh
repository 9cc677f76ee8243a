#!/bin/bash
L=${1:-r2r}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph.py tests/test_gpu_attn_dropout.py tests/test_gpu_training.py tests/test_gpu_layernorm.py tests/test_gpu_mlp.py -q > gpurun_out/${L}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${L}_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=|Error" gpurun_out/${L}_pytest.log | tail -n 12 | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --drop-rate 0.1 > gpurun_out/${L}_bench_drop01.log 2>&1; echo "bench rc=$?" >> gpurun_out/${L}_bench_drop01.log
tail -n 3 gpurun_out/${L}_bench_drop01.log | cut -c1-500
timeout 800 python scripts/configs_check.py > gpurun_out/${L}_all_configs.log 2>&1
tail -n 8 gpurun_out/${L}_all_configs.log | cut -c1-200
