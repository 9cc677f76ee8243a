#!/bin/bash
# final round-1 verification: GPU tests, smoke, bench (ours + reference arm + reference-default drop rates), step
# composition, all BASELINE configs, kernel micro-timings, ncu captures of the streaming kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1q_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1q_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r1q_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-budget-s 60 > gpurun_out/r1q_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r1q_bench.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --drop-rate 0.1 > gpurun_out/r1q_bench_drop01.log 2>&1
timeout 400 python scripts/gap_check.py > gpurun_out/r1q_step_composition.log 2>&1
timeout 800 python scripts/configs_check.py > gpurun_out/r1q_all_configs.log 2>&1
timeout 300 python scripts/tc_check.py --time > gpurun_out/r1q_tc_check.log 2>&1
timeout 300 python scripts/ln_check.py > gpurun_out/r1q_ln_check.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ln_fwd|ln_bwd|bias_gelu" -s 6 -c 4 -o gpurun_out/r1q_stream python scripts/gap_check.py > gpurun_out/r1q_ncu_stream.log 2>&1
tail -n 3 gpurun_out/r1q_pytest.log gpurun_out/r1q_smoke.log; cut -c1-260 gpurun_out/r1q_bench.log; cut -c1-260 gpurun_out/r1q_bench_drop01.log; tail -n 6 gpurun_out/r1q_all_configs.log
