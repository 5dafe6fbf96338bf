#!/bin/bash
# usage: tools/gpu_bench.sh <tag> [pytest|nopytest] [env...]: optional GPU test suite, then one bench line (main workload only)
tag=$1; shift
mode=${1:-nopytest}; shift
mkdir -p gpurun_out
if [ "$mode" = "pytest" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
  tail -n 5 gpurun_out/${tag}_pytest.log
fi
export OSC2_BENCH_EXTRAS=0 OSC2_BENCH_CASE2=0 "$@"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.log 2>gpurun_out/${tag}_bench.err
python - gpurun_out/${tag}_bench.log <<'PY'
import json,sys
f=sys.argv[1]
try:
    l=[x for x in open(f) if x.startswith("{")][-1]; j=json.loads(l)
    print(f, "%.4g"%j["value"], "%.2f ms"%j["ms_per_step"], {k:round(v,2) for k,v in j["roofline"]["step"]["kernel_ms"].items()}, "step frac %.3f"%j["roofline"]["step"]["frac"], "parity", j["parity_check"]["ok"], j["parity_check"]["max_rel_velocity"])
except Exception as e:
    print(f, "ERR", e); print(open(f).read()[-1500:]); print(open(f.replace(".log",".err")).read()[-3000:])
PY
