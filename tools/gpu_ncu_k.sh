#!/bin/bash
# usage: tools/gpu_ncu_k.sh <tag> <kernel regex> <skip> <count> [env...]: one ncu --set full capture of the matching kernels
tag=$1; rx=$2; skip=$3; cnt=$4; shift; shift; shift; shift
mkdir -p gpurun_out
export OSC2_BENCH_CASE2=0 OSC2_BENCH_EXTRAS=0 "$@"
ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt \
    -o gpurun_out/${tag} python bench.py --steps 2 --warmup 3 > gpurun_out/${tag}_ncu.log 2>&1
echo "rc=$?"; ls -la gpurun_out/${tag}.ncu-rep
