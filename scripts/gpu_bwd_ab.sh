#!/bin/bash
# A/B of attention-backward build variants (tools/trace_bwd_<variant>, built with -DHS_BWD_* switches) on one box:
# stage-0 time and unit period, cos and non-cos
mkdir -p gpurun_out
for rep in 1 2; do
for v in "$@"; do
  for cos in 1 0; do
    ./tools/trace_bwd_$v $cos > gpurun_out/ab_$v.log 2>&1
    echo "$v cos=$cos: $(grep 'run 2' gpurun_out/ab_$v.log | cut -d' ' -f6-7) $(grep 'unit period' gpurun_out/ab_$v.log)"
  done
done
done
