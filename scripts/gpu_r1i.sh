#!/bin/bash
# r1i: full GPU test suite, smoke, bench (TF32 and fp32 library GEMMs), ncu launch list, full ncu captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1i_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1i_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r1i_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-budget-s 60 > gpurun_out/r1i_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r1i_bench.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --gemm-precision fp32 > gpurun_out/r1i_bench_fp32gemm.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 --cpu-budget-s 60 > gpurun_out/r1i_bench_reference.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r1i_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1i_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd_tc|attn_fwd_tc" -s 36 -c 4 -o gpurun_out/r1i_attn python scripts/tc_check.py --time > gpurun_out/r1i_ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ln_fwd|ln_bwd" -s 12 -c 2 -o gpurun_out/r1i_ln python scripts/ln_check.py > gpurun_out/r1i_ncu_ln.log 2>&1
tail -n 3 gpurun_out/r1i_pytest.log gpurun_out/r1i_smoke.log; cut -c1-300 gpurun_out/r1i_bench.log; cut -c1-300 gpurun_out/r1i_bench_fp32gemm.log; cut -c1-300 gpurun_out/r1i_bench_reference.log
