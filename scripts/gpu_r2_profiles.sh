#!/bin/bash
# round-2 evidence run: ncu launch list of two eager bench steps, ncu --set full captures of the GEMM in its regimes and of
# the attention kernels inside a bench step, GEMM micro-timings
L=${1:-r2}
mkdir -p gpurun_out
timeout 900 python scripts/gemm3_check.py --time > gpurun_out/${L}_gemm3_check.log 2>&1; echo "rc=$?" >> gpurun_out/${L}_gemm3_check.log
T0=1572864
for spec in "qkv $T0 288 96 0 0" "fc1gelu $T0 384 96 2 0" "fc2 $T0 96 384 0 0" "s2fc1 98304 1536 384 0 0" "s2dgrad_tf32 98304 384 1536 1 1"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm3 -s 2 -c 1 -f -o gpurun_out/${L}_g3_$1 \
    python scripts/gemm3_prof.py $2 $3 $4 $5 $6 > gpurun_out/${L}_ncu_$1.log 2>&1
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${L}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cuda-graph > gpurun_out/${L}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:attn_.*_tc_kernel -s 2 -c 2 -f -o gpurun_out/${L}_attn \
  python scripts/tc_check.py --time > gpurun_out/${L}_ncu_attn.log 2>&1
ls -la gpurun_out/${L}_*.ncu-rep gpurun_out/${L}_launches.csv
