#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/dp2_gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/dp2_bench.log 2>&1; echo "rc=$?" >> gpurun_out/dp2_bench.log
tail -n 4 gpurun_out/dp2_bench.log | cut -c1-1200
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/dp1_bench.log 2>&1; echo "rc=$?" >> gpurun_out/dp1_bench.log
tail -n 2 gpurun_out/dp1_bench.log | cut -c1-3000
