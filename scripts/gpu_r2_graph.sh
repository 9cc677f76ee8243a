#!/bin/bash
L=${1:-r2f}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_graph.py tests/test_gpu_attention.py tests/test_gpu_model.py -q > gpurun_out/${L}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${L}_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=|Error" gpurun_out/${L}_pytest.log | tail -n 20
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-budget-s 30 > gpurun_out/${L}_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/${L}_bench.log
tail -n 12 gpurun_out/${L}_bench.log | cut -c1-700
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-cuda-graph > gpurun_out/${L}_bench_eager.log 2>&1; echo "bench rc=$?" >> gpurun_out/${L}_bench_eager.log
tail -n 3 gpurun_out/${L}_bench_eager.log | cut -c1-400
