#!/bin/bash
# usage: tools/gpu_ncu_all.sh <tag> [env...]: plain bench line, then one ncu --set full capture of the four main kernels of one step
tag=$1; shift
mkdir -p gpurun_out
export OSC2_BENCH_CASE2=0 OSC2_BENCH_EXTRAS=0 "$@"
python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_plain.log 2>gpurun_out/${tag}_plain.err || { echo "bench failed"; tail -20 gpurun_out/${tag}_plain.err; exit 1; }
tail -c 600 gpurun_out/${tag}_plain.log
ncu --set full --clock-control none --import-source on -k regex:"osc2_(forward|backward|gauss|opcond)" -s 12 -c 4 \
    -o gpurun_out/${tag} python bench.py --steps 2 --warmup 3 > gpurun_out/${tag}_ncu.log 2>&1
echo "rc=$?"; ls -la gpurun_out/${tag}.ncu-rep
