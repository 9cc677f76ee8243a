#!/bin/bash
# attention kernels: tensor-core vs exact check + stage-0 timings, then the attention test files
L=${1:-attn}
mkdir -p gpurun_out
timeout 600 python scripts/tc_check.py --time > gpurun_out/${L}_tc_check.log 2>&1; echo "tc_check rc=$?" >> gpurun_out/${L}_tc_check.log
grep -E "stage-0|passed|FAIL|rc=|Error|error" gpurun_out/${L}_tc_check.log | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_attention_tc.py tests/test_gpu_attention.py tests/test_gpu_attn_dropout.py tests/test_gpu_model.py -m gpu -q -x > gpurun_out/${L}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${L}_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/${L}_pytest.log | tail -n 8 | cut -c1-300
