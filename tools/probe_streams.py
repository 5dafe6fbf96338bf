"""Experiment: split the batch over K handles (K streams) on one GPU so that the latency-bound forward
sweep of one sub-batch overlaps the throughput kernels of the others."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from opensc2_b200.batch import BatchedStep

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = 5
B, N = bench.BATCH, bench.NODES
sub = B // K
hs = []
for k in range(K):
    topo, model, tab, arr, sweep = bench.make_workload(sub, N, seed=100 + k)
    bs = BatchedStep(topo, sub, N)
    bs.set_model(model, [tab]); bs.upload_sweep(sweep); bs.upload_state(arr.sysvar, None); bs.upload_geometry(arr); bs.synchronize()
    hs.append(bs)
n = 0
for _ in range(3):
    n += 1
    for bs in hs: bs.advance(bench.T_START + n * bench.DT, bench.DT, 1 + n)
for bs in hs: bs.synchronize()
t0 = time.perf_counter()
for bs in hs: bs.timer_start()
for _ in range(steps):
    n += 1
    for bs in hs: bs.advance(bench.T_START + n * bench.DT, bench.DT, 1 + n)
ms = [bs.timer_stop() for bs in hs]
wall = (time.perf_counter() - t0) * 1e3
print(f"K={K} sub-batch={sub}: device ms per step (max over streams) {max(ms)/steps:.2f}, wall {wall/steps:.2f}, "
      f"node-timesteps/s {K*sub*N*steps/(max(ms)*1e-3):.4g}")
