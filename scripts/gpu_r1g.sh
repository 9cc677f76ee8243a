#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1g_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1g_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r1g_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r1g_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1g_ncu_bench.log 2>&1
tail -n 3 gpurun_out/r1g_pytest.log; cut -c1-330 gpurun_out/r1g_bench.log
