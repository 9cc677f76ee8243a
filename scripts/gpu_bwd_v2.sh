#!/bin/bash
# r2-wip: first contact of the 16-warp attention backward (HEALSWIN_ATTN_BWD_V2=1) with the hardware:
# parity tests of the tensor-core attention under a short timeout (a protocol bug traps after ~2 s instead of hanging),
# then isolated timings of both variants.
mkdir -p gpurun_out
HEALSWIN_ATTN_BWD_V2=1 timeout 300 python -m pytest tests/test_gpu_attention_tc.py tests/test_gpu_attn_dropout.py -x -q > gpurun_out/bwd_v2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/bwd_v2_pytest.log
timeout 200 python scripts/tc_check.py --time > gpurun_out/bwd_v1_time.log 2>&1
HEALSWIN_ATTN_BWD_V2=1 timeout 200 python scripts/tc_check.py --time > gpurun_out/bwd_v2_time.log 2>&1
tail -n 5 gpurun_out/bwd_v2_pytest.log; tail -n 4 gpurun_out/bwd_v1_time.log gpurun_out/bwd_v2_time.log
