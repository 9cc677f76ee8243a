#!/bin/bash
# r1d: tensor-core forward + backward: GPU tests, smoke, bench, ncu launch list, full ncu capture of the TC bwd kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1d_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1d_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r1d_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-budget-s 40 > gpurun_out/r1d_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r1d_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r1d_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1d_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc -s 9 -c 1 -o gpurun_out/r1d_attn_bwd_tc python scripts/tc_check.py --time > gpurun_out/r1d_ncu_full.log 2>&1
tail -n 3 gpurun_out/r1d_pytest.log gpurun_out/r1d_smoke.log gpurun_out/r1d_bench.log
