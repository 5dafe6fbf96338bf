#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_6.log
tail -n 4 gpurun_out/r2_pytest_gpu_6.log
show() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    l=[x for x in open(f) if x.startswith("{")][-1]; j=json.loads(l)
    print(f, "%.4g"%j["value"], "%.2f ms"%j["ms_per_step"], {k:round(v,2) for k,v in j["roofline"]["step"]["kernel_ms"].items()}, "frac %.3f"%j["roofline"]["frac"], j["roofline"]["kernel"], j["parity_check"]["max_rel_velocity"])
except Exception as e: print(f, "ERR", e); print(open(f).read()[-1500:])
PY
}
OSC2_BENCH_EXTRAS=0 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_6_fused.log 2>&1; show gpurun_out/r2_bench_6_fused.log
OSC2_FUSED_PG=0 OSC2_BENCH_EXTRAS=0 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_6_split.log 2>&1; show gpurun_out/r2_bench_6_split.log
