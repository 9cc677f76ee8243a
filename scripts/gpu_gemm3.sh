#!/bin/bash
# bf16x3 GEMM: tiny shapes, every linear-layer shape of the bench model with timings, ncu captures of three regimes
L=${1:-r2a}
mkdir -p gpurun_out
timeout 180 python scripts/gemm3_check.py --quick > gpurun_out/${L}_gemm3_quick.log 2>&1; echo "quick rc=$?" >> gpurun_out/${L}_gemm3_quick.log
tail -n 4 gpurun_out/${L}_gemm3_quick.log
timeout 900 python scripts/gemm3_check.py --time > gpurun_out/${L}_gemm3_check.log 2>&1; echo "full rc=$?" >> gpurun_out/${L}_gemm3_check.log
grep -E "^ +plain|FAIL|rc=|gemm3_check" gpurun_out/${L}_gemm3_check.log
if [ "$2" == "ncu" ]; then
  T0=1572864
  for spec in "qkv $T0 288 96 0" "fc2 $T0 96 384 0" "s2fc1 98304 1536 384 0" "fc1gelu $T0 384 96 2"; do
    set -- $spec
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm3 -s 2 -c 1 -f -o gpurun_out/${L}_g3_$1 \
      python scripts/gemm3_prof.py $2 $3 $4 $5 > gpurun_out/${L}_ncu_$1.log 2>&1
  done
  ls -la gpurun_out/*.ncu-rep
fi
