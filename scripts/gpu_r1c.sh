#!/bin/bash
# r1c: tensor-core forward path: GPU tests, smoke, bench, ncu launch list, one full ncu capture of the TC fwd kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r1c_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1c_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1c_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r1c_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-budget-s 40 > gpurun_out/r1c_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r1c_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r1c_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1c_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 10 -c 2 -o gpurun_out/r1c_attn_fwd_tc python scripts/tc_check.py --time > gpurun_out/r1c_ncu_full.log 2>&1
tail -3 gpurun_out/r1c_pytest.log gpurun_out/r1c_smoke.log gpurun_out/r1c_bench.log
