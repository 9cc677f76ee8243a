#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/tc_check.py --time > gpurun_out/tc_check.log 2>&1; echo "rc=$?" >> gpurun_out/tc_check.log
cat gpurun_out/tc_check.log
if grep -q "tc_check passed" gpurun_out/tc_check.log; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tc_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tc_pytest.log
  tail -n 15 gpurun_out/tc_pytest.log
fi
