#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_2.log
tail -n 6 gpurun_out/r2_pytest_gpu_2.log
OSC2_BENCH_CASE2=0 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_2_tile.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_2_tile.log",):
    try:
        l=[x for x in open(f) if x.startswith("{")][-1]; j=json.loads(l)
        print(f, j["value"], j["ms_per_step"], j["roofline"]["step"]["kernel_ms"], j["roofline"]["frac"])
    except Exception as e: print(f, "ERR", e); print(open(f).read()[-2000:])
PY
OSC2_FWD_TILE=0 OSC2_BENCH_CASE2=0 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_2_old.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_2_old.log",):
    try:
        l=[x for x in open(f) if x.startswith("{")][-1]; j=json.loads(l)
        print(f, j["value"], j["ms_per_step"], j["roofline"]["step"]["kernel_ms"], j["roofline"]["frac"])
    except Exception as e: print(f, "ERR", e); print(open(f).read()[-2000:])
PY
