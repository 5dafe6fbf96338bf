"""Quick device-resident timing of the batched step kernels (development probe)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

from opensc2_b200.batch import BatchedStep
from opensc2_b200.problem import case1_topology, case2_topology, case3_topology, synthetic_step_arrays

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--nodes", type=int, default=1001)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--case", default="case1")
ap.add_argument("--tile", type=int, default=64, help="distinct conductors generated on the host, tiled to the batch")
a = ap.parse_args()
topo = {"case1": case1_topology(1), "case2": case2_topology(), "case3": case3_topology()}[a.case]
t0 = time.time()
small = synthetic_step_arrays(topo, a.nodes, batch=min(a.tile, a.batch), seed=1)
import dataclasses
kw = {}
reps = (a.batch + small.dz.shape[0] - 1) // small.dz.shape[0]
for f in dataclasses.fields(small):
    v = getattr(small, f.name)
    kw[f.name] = np.concatenate([v] * reps, axis=0)[:a.batch] if isinstance(v, np.ndarray) else v
arr = type(small)(**kw)
print("host arrays %.1fs" % (time.time() - t0), flush=True)
with BatchedStep(topo, a.batch, a.nodes) as bs:
    print("device GB", bs.device_bytes / 1e9, flush=True)
    t0 = time.time()
    bs.upload_state(arr.sysvar, arr.syslod)
    bs.upload_step_inputs(arr)
    bs.synchronize()
    print("upload %.2fs" % (time.time() - t0), flush=True)
    for k in range(2):
        bs.step(arr.dt, 2 + k)
    bs.synchronize()
    bs.set_profiling(True)
    bs.kernel_ms(reset=True)
    bs.timer_start()
    for k in range(a.steps):
        bs.step(arr.dt, 4 + k)
    ms = bs.timer_stop()
    km = bs.kernel_ms()
    bs.synchronize()
    nts = a.batch * a.nodes * a.steps / (ms * 1e-3)
    d = topo.ndf
    balg = 8 * (3 * d * (4 * d - 1) + 6 * d)
    print(json.dumps({"batch": a.batch, "nodes": a.nodes, "steps": a.steps, "ms_per_step": ms / a.steps,
                      "node_timesteps_per_s": nts, "alg_GBps": nts * balg / 1e9,
                      "kernel_ms_per_step": {k: v[0] / a.steps for k, v in km.items()}}))
