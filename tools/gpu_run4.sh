#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_4.log 2> gpurun_out/r2_bench_4.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench_4.err
python - <<'PY'
import json
l=[x for x in open("gpurun_out/r2_bench_4.log") if x.startswith("{")]
if l:
    j=json.loads(l[-1])
    for k in ("value","ms_per_step","parity_check","snapshot","e2e","e2e_host_step","cpu_baseline"): print(k, j.get(k))
    print(j["roofline"]["step"]["kernel_ms"], j["roofline"]["frac"], j["roofline"]["traffic"], j["roofline"]["traffic_source"])
    print(json.dumps(j.get("other_workloads"))[:1500])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_4_ref.log 2>&1; tail -c 600 gpurun_out/r2_bench_4_ref.log
