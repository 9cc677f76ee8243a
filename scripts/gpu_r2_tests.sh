#!/bin/bash
# full GPU test-suite (all failures listed) + first contact of the 16-warp attention backward
L=${1:-r2e}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${L}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${L}_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/${L}_pytest.log | tail -n 30
HEALSWIN_ATTN_BWD_V2=1 timeout 300 python -m pytest tests/test_gpu_attention_tc.py tests/test_gpu_attn_dropout.py -x -q > gpurun_out/${L}_bwd_v2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${L}_bwd_v2_pytest.log
timeout 200 python scripts/tc_check.py --time > gpurun_out/${L}_bwd_v1_time.log 2>&1
HEALSWIN_ATTN_BWD_V2=1 timeout 200 python scripts/tc_check.py --time > gpurun_out/${L}_bwd_v2_time.log 2>&1
tail -n 8 gpurun_out/${L}_bwd_v2_pytest.log; tail -n 5 gpurun_out/${L}_bwd_v1_time.log gpurun_out/${L}_bwd_v2_time.log
