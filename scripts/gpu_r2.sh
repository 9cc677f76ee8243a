#!/bin/bash
# round-2 verification run (label $1): GPU tests, smoke, bench, step composition; optional extras by $2
L=${1:-r2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${L}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${L}_pytest.log
tail -n 15 gpurun_out/${L}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${L}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${L}_smoke.log
tail -n 3 gpurun_out/${L}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-budget-s 30 > gpurun_out/${L}_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/${L}_bench.log
cut -c1-1500 gpurun_out/${L}_bench.log | tail -n 5
timeout 400 python scripts/gap_check.py > gpurun_out/${L}_step_composition.log 2>&1
head -n 40 gpurun_out/${L}_step_composition.log
if [ "$2" == "all" ]; then
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --drop-rate 0.1 > gpurun_out/${L}_bench_drop01.log 2>&1
  timeout 800 python scripts/configs_check.py > gpurun_out/${L}_all_configs.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${L}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${L}_ncu_bench.log 2>&1
  cut -c1-200 gpurun_out/${L}_bench_drop01.log; tail -n 6 gpurun_out/${L}_all_configs.log
fi
