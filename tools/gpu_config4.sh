#!/bin/bash
# usage: tools/gpu_config4.sh N   -- BASELINE configs[4] on N GPUs (strong scaling of the fixed 131 072-variant job)
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --workload config4 --gpus 1 --steps 3 --warmup 3 > gpurun_out/r2_config4_n1.log 2> gpurun_out/r2_config4_n1.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --workload config4 --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_config4_n$N.log 2> gpurun_out/r2_config4_n$N.err
fi
echo "rc=$?"; tail -c 600 gpurun_out/r2_config4_n$N.err; grep '^{' gpurun_out/r2_config4_n$N.log | tail -n 1 | cut -c1-400
