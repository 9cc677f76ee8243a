#!/bin/bash
# GEMM check (quick [+ full timing with $2 == time]), the GPU test-suite, bench (graph replay)
L=${1:-r2h}
mkdir -p gpurun_out
timeout 180 python scripts/gemm3_check.py --quick > gpurun_out/${L}_gemm3_quick.log 2>&1; echo "quick rc=$?" >> gpurun_out/${L}_gemm3_quick.log
tail -n 2 gpurun_out/${L}_gemm3_quick.log | cut -c1-300
if grep -q "ALL OK" gpurun_out/${L}_gemm3_quick.log; then
  if [ "$2" == "time" ]; then
    timeout 900 python scripts/gemm3_check.py --time > gpurun_out/${L}_gemm3_check.log 2>&1; echo "full rc=$?" >> gpurun_out/${L}_gemm3_check.log
    grep -E "^ +plain|FAIL|rc=|gemm3_check" gpurun_out/${L}_gemm3_check.log | cut -c1-260
  fi
  timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${L}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${L}_pytest.log
  grep -E "^(FAILED|ERROR)|passed|failed|rc=|Error" gpurun_out/${L}_pytest.log | tail -n 20 | cut -c1-300
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${L}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${L}_smoke.log
  tail -n 2 gpurun_out/${L}_smoke.log
  timeout 600 python bench.py --steps 10 --warmup 3 --cpu-budget-s 30 > gpurun_out/${L}_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/${L}_bench.log
  tail -n 4 gpurun_out/${L}_bench.log | cut -c1-400
fi
