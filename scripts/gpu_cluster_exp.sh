#!/bin/bash
mkdir -p gpurun_out
for wr in 3 4 5; do
  echo "== W ring $wr"
  HEALSWIN_GEMM3_WRING=$wr timeout 300 python scripts/gemm3_check.py --time --tf32-dgrad --only=s2 --only=s3 --only=s1 2>&1 | grep -E "^ +plain|FAIL|ALL" | cut -c1-250
done > gpurun_out/r2q_wring_exp.log 2>&1
cat gpurun_out/r2q_wring_exp.log
