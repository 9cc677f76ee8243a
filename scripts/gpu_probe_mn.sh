#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/probe_mn.log
for v in 0 1 2 3 4 5; do timeout 30 ./tools/probe_mn $v >> gpurun_out/probe_mn.log 2>&1; echo "rc=$?" >> gpurun_out/probe_mn.log; done
cat gpurun_out/probe_mn.log
