#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log
tail -n 25 gpurun_out/q_pytest.log
timeout 300 python scripts/tc_check.py --time 2>&1 | grep -E "stage-0|tc_check"
