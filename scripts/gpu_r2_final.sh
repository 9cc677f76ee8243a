#!/bin/bash
# round-2 closing evidence: ncu --set full of the attention backward at the BASELINE stage-0 shape (cos + bias), its phase
# trace, the ncu launch list of two eager bench steps, then checks + the whole GPU suite + smoke + bench (graph replay)
L=${1:-r2f}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc_kernel -s 1 -c 1 -f -o gpurun_out/${L}_attn_bwd ./tools/bwd_run 1 > gpurun_out/${L}_ncu_attn_bwd.log 2>&1
timeout 120 ./tools/trace_bwd 1 > gpurun_out/${L}_trace_bwd.log 2>&1; tail -n 8 gpurun_out/${L}_trace_bwd.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${L}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cuda-graph > gpurun_out/${L}_ncu_bench.log 2>&1
timeout 300 python scripts/gap_check.py > gpurun_out/${L}_gap.log 2>&1; head -n 14 gpurun_out/${L}_gap.log | cut -c1-140
bash scripts/gpu_r2_combo.sh ${L}
