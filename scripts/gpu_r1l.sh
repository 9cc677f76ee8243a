#!/bin/bash
# r1l: final round-1 evaluation: tests, smoke, bench (+ reference arm), ncu launch list, ncu full captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1l_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1l_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r1l_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-budget-s 60 > gpurun_out/r1l_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r1l_bench.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 --cpu-budget-s 60 > gpurun_out/r1l_bench_reference.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4500 --csv --log-file gpurun_out/r1l_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1l_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc -s 9 -c 1 -o gpurun_out/r1l_attn_bwd python scripts/tc_check.py --time > gpurun_out/r1l_ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc -s 4 -c 2 -o gpurun_out/r1l_wgrad python scripts/wgrad_check.py > gpurun_out/r1l_ncu_wgrad.log 2>&1
tail -n 3 gpurun_out/r1l_pytest.log gpurun_out/r1l_smoke.log; cut -c1-300 gpurun_out/r1l_bench.log; cut -c1-200 gpurun_out/r1l_bench_reference.log
