#!/bin/bash
# cta_group::2 ("pair") GEMM: parity + timings of the tensor-bound shapes, pair off / on (and extra env), on one box
L=${1:-pair}
mkdir -p gpurun_out
run() {  # label, env...
  local tag=$1; shift
  env "$@" timeout 400 python scripts/gemm3_check.py --time --tf32-dgrad --only="s2 fc1" --only="s3 fc2" > gpurun_out/${L}_$tag.log 2>&1
  echo "$tag rc=$?" >> gpurun_out/${L}_$tag.log
  grep -E "FAIL|plain|rc=|rror" gpurun_out/${L}_$tag.log | cut -c1-250
}
run p0 HEALSWIN_GEMM3_PAIR=0
run p1 HEALSWIN_GEMM3_PAIR=1
run p1w4 HEALSWIN_GEMM3_PAIR=1 HEALSWIN_GEMM3_WRING=4
run p1w6 HEALSWIN_GEMM3_PAIR=1 HEALSWIN_GEMM3_WRING=6
