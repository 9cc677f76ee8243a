#!/bin/bash
# round-1 verification run (label $1, default r1r): GPU tests, smoke, bench (ours + reference-default drop rates), step
# composition, all BASELINE configs, kernel micro-timings, ncu launch list of two bench steps
L=${1:-r1r}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${L}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${L}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${L}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${L}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-budget-s 30 > gpurun_out/${L}_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/${L}_bench.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --drop-rate 0.1 > gpurun_out/${L}_bench_drop01.log 2>&1
timeout 400 python scripts/gap_check.py > gpurun_out/${L}_step_composition.log 2>&1
timeout 800 python scripts/configs_check.py > gpurun_out/${L}_all_configs.log 2>&1
timeout 200 python scripts/mlp_check.py > gpurun_out/${L}_mlp_check.log 2>&1
timeout 200 python scripts/ln_head_check.py > gpurun_out/${L}_ln_head_check.log 2>&1
timeout 200 python scripts/dgrad_acc_check.py > gpurun_out/${L}_dgrad_acc_check.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${L}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${L}_ncu_bench.log 2>&1
tail -n 3 gpurun_out/${L}_pytest.log gpurun_out/${L}_smoke.log; cut -c1-300 gpurun_out/${L}_bench.log; cut -c1-200 gpurun_out/${L}_bench_drop01.log; tail -n 6 gpurun_out/${L}_all_configs.log; head -n 14 gpurun_out/${L}_step_composition.log
