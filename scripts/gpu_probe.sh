#!/bin/bash
# runs every probe variant in its own process (a faulting variant poisons its CUDA context)
mkdir -p gpurun_out
: > gpurun_out/probe.log
for v in 1 2 3 40 41 42 43 50 51 52 9; do
  echo "=== variant $v" >> gpurun_out/probe.log
  timeout 60 ./tools/probe_umma $v >> gpurun_out/probe.log 2>&1
  echo "rc=$?" >> gpurun_out/probe.log
done
cat gpurun_out/probe.log
