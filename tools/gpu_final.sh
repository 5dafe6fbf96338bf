#!/bin/bash
# usage: tools/gpu_final.sh <tag>: GPU test suite, the default bench line (with the extras), the reference arm, then the profile
tag=$1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log; tail -n 3 gpurun_out/${tag}_pytest.log
timeout 1200 python bench.py > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.log 2> gpurun_out/${tag}_bench_reference.err; echo "reference rc=$?"
tools/gpu_profile_final.sh ${tag}_prof
