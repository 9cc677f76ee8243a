#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/trace_bwd 1 > gpurun_out/trace_bwd.log 2>&1; tail -n 8 gpurun_out/trace_bwd.log
timeout 300 python scripts/tc_check.py --time > gpurun_out/tc_check.log 2>&1; echo "rc=$?" >> gpurun_out/tc_check.log
grep -E "BWD|stage-0|tc_check|rel " gpurun_out/tc_check.log
if grep -q "tc_check passed" gpurun_out/tc_check.log; then
  timeout 600 python -m pytest tests/test_gpu_attention_tc.py tests/test_gpu_model.py -x -q > gpurun_out/tc_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tc_pytest.log
  tail -n 4 gpurun_out/tc_pytest.log
fi
