#!/bin/bash
# final round-1 verification: GPU tests, smoke, bench (ours + reference arm), step composition
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1n_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1n_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r1n_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-budget-s 60 > gpurun_out/r1n_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r1n_bench.log
timeout 400 python scripts/gap_check.py > gpurun_out/r1n_step_composition.log 2>&1
timeout 300 python scripts/tc_check.py --time > gpurun_out/r1n_tc_check.log 2>&1
timeout 300 python scripts/ln_check.py > gpurun_out/r1n_ln_check.log 2>&1
tail -n 3 gpurun_out/r1n_pytest.log gpurun_out/r1n_smoke.log; cut -c1-300 gpurun_out/r1n_bench.log
