#!/bin/bash
# usage: tools/gpu_ncu2.sh <tag> <kernel regex> [env...]
tag=$1; rx=$2; shift; shift
mkdir -p gpurun_out
export OSC2_BENCH_CASE2=0 "$@"
python bench.py --steps 2 --warmup 3 > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$rx" -s 3 -c 1 \
    -o gpurun_out/${tag} python bench.py --steps 2 --warmup 3 > gpurun_out/${tag}_ncu.log 2>&1
echo "rc=$?"; ls -la gpurun_out/${tag}.ncu-rep
