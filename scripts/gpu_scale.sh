#!/bin/bash
# weak-scaling run at N GPUs (argument), as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/scale_${N}_gpus.txt
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${N}.log 2>&1
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_${N}.log 2>&1
fi
echo "rc=$?" >> gpurun_out/scale_${N}.log
tail -n 2 gpurun_out/scale_${N}.log | cut -c1-420
