#!/bin/bash
# usage: tools/gpu_ncu.sh <tag> [env assignments...]   -- ncu --set full of the sweep + assembly kernels at bench size
tag=$1; shift
mkdir -p gpurun_out
export OSC2_BENCH_CASE2=0 "$@"
python bench.py --steps 2 --warmup 3 > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'osc2_forward|osc2_gauss2|osc2_backward|osc2_opcond' -s 8 -c 4 \
    -o gpurun_out/${tag} python bench.py --steps 2 --warmup 3 > gpurun_out/${tag}_ncu.log 2>&1
echo "rc=$?"; tail -n 3 gpurun_out/${tag}_ncu.log; ls -la gpurun_out/${tag}.ncu-rep
