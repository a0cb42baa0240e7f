#!/bin/bash
# usage: run_scaling.sh N [extra env]   -- bench at N GPUs (torchrun for N > 1)
N=$1; shift
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  env "$@" python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/scale_$N.err | tail -1
else
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/scale_$N.err | tail -1
fi
