#!/bin/bash
# GPU call 1 of round 2: the whole GPU suite (incl. the transient pins) + a baseline bench line.
mkdir -p gpurun_out
export OSC2_TRANSIENT_REPORT=gpurun_out/r2_transient_drift.tsv
rm -f $OSC2_TRANSIENT_REPORT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_1.log
tail -n 15 gpurun_out/r2_pytest_gpu_1.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_1.log 2>&1; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2_bench_1.log
