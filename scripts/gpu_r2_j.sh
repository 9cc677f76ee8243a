#!/bin/bash
L=${1:-r2j}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm3.py tests/test_gpu_model.py tests/test_flat_model.py tests/test_gpu_graph.py -q > gpurun_out/${L}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${L}_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=|Error" gpurun_out/${L}_pytest.log | tail -n 20 | cut -c1-300
timeout 400 python scripts/gap_check.py > gpurun_out/${L}_step_composition.log 2>&1
head -n 48 gpurun_out/${L}_step_composition.log | cut -c1-170
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${L}_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/${L}_bench.log
tail -n 3 gpurun_out/${L}_bench.log | cut -c1-300
