"""Throughput probe of the NODOFS = 64 path: CASE_2-sized conductors, device-resident full steps
(the same sample bench.py reports under `other_workloads`, on the deck's 151-node mesh).   usage: probe_case2.py [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = 151
r = bench.case2_sample(0, batch=batch)
print(f"CASE_2 d=64: {batch} x {n} nodes, {r['ms_per_step']:.2f} ms/step, {r['value']:.4g} node-timesteps/s",
      {k: round(v, 2) for k, v in r["kernel_ms"].items()})
