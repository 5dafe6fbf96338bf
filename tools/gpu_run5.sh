#!/bin/bash
mkdir -p gpurun_out
OSC2_BENCH_EXTRAS=0 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_5.log 2> gpurun_out/r2_bench_5.err; echo "bench rc=$?"
python - <<'PY'
import json
l=[x for x in open("gpurun_out/r2_bench_5.log") if x.startswith("{")]
j=json.loads(l[-1]); print("value", j["value"], "e2e", j["e2e"]["value"], j["e2e"]["ms_per_step"], j["parity_check"]["max_rel_velocity"])
PY
OSC2_BENCH_TOTAL4=16384 timeout 900 python bench.py --workload config4 --steps 3 --warmup 3 > gpurun_out/r2_config4_small.log 2> gpurun_out/r2_config4_small.err; echo "config4 rc=$?"
tail -c 800 gpurun_out/r2_config4_small.err; tail -c 1800 gpurun_out/r2_config4_small.log
