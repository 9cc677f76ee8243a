#!/bin/bash
# A/B: TF32 truncation compensation on/off (errors vs the exact fp32 kernels), then GPU tests
mkdir -p gpurun_out
echo "=== compensation ON" > gpurun_out/ab.log
timeout 300 python scripts/tc_check.py >> gpurun_out/ab.log 2>&1
echo "=== compensation OFF" >> gpurun_out/ab.log
HEALSWIN_NO_TRUNC_COMP=1 timeout 300 python scripts/tc_check.py >> gpurun_out/ab.log 2>&1
cat gpurun_out/ab.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ab_pytest.log
tail -n 15 gpurun_out/ab_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
