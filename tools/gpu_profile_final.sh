#!/bin/bash
# usage: tools/gpu_profile_final.sh <tag>: plain bench line, launch list (gpu__time_duration per launch), one ncu --set full capture
# of the six kernels of one step of the main workload
tag=$1; shift
mkdir -p gpurun_out
export OSC2_BENCH_CASE2=0 OSC2_BENCH_EXTRAS=0 "$@"
python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_plain.log 2>gpurun_out/${tag}_plain.err || { echo "bench failed"; tail -20 gpurun_out/${tag}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/${tag}_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"osc2_(forward|backward|gauss|opcond)" -s 18 -c 6 \
    -o gpurun_out/${tag} python bench.py --steps 2 --warmup 3 > gpurun_out/${tag}_ncu.log 2>&1
echo "rc=$?"; ls -la gpurun_out/${tag}.ncu-rep
