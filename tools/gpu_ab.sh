#!/bin/bash
# usage: tools/gpu_ab.sh <tag> <pytest-args or -> <ENV=val for arm B>: pytest once (arm A = defaults), bench arm A, bench arm B
tag=$1; pt=$2; shift; shift
mkdir -p gpurun_out
if [ "$pt" != "-" ]; then
  timeout 1500 python -m pytest $pt -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
  tail -n 5 gpurun_out/${tag}_pytest.log
fi
tools/gpu_bench.sh ${tag}_A nopytest
tools/gpu_bench.sh ${tag}_B nopytest "$@"
