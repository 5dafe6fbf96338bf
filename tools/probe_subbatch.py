"""Does splitting the batch over several handles (one stream each) pay?  The sweeps are latency
bound (every conductor advances one node per ~9000 cycles whatever the batch size), so sub-batches
whose steps drift apart can fill each other's idle issue slots and HBM bandwidth."""
import argparse
import dataclasses
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
from opensc2_b200.batch import BatchedStep

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--nodes", type=int, default=2001)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--subs", type=int, nargs="+", default=[1, 2, 4, 8])
ap.add_argument("--stagger", type=int, default=0, help="extra untimed steps given to sub-batch q: q * stagger / subs of a step is not possible, so whole warm-up steps")
a = ap.parse_args()
topo, model, tab, arr, sweep = bench.make_workload(a.batch, a.nodes, seed=1)


def sl(obj, lo, hi):
    kw = {}
    for f in dataclasses.fields(obj):
        v = getattr(obj, f.name)
        kw[f.name] = np.ascontiguousarray(v[lo:hi]) if isinstance(v, np.ndarray) and v.shape[0] == a.batch else v
    return type(obj)(**kw)


for S in a.subs:
    nb = a.batch // S
    hs = []
    for q in range(S):
        bs = BatchedStep(topo, nb, a.nodes)
        bs.set_model(model, [tab])
        bs.upload_sweep(sl(sweep, q * nb, (q + 1) * nb))
        sub = sl(arr, q * nb, (q + 1) * nb)
        bs.upload_state(sub.sysvar, None)
        bs.upload_geometry(sub)
        hs.append(bs)
    for bs in hs:
        bs.synchronize()
    k = 0
    for _ in range(2):
        k += 1
        for bs in hs:
            bs.advance(bench.T_START + k * bench.DT, bench.DT, k)
    for bs in hs:
        bs.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        k += 1
        for bs in hs:
            bs.advance(bench.T_START + k * bench.DT, bench.DT, k)
    for bs in hs:
        bs.synchronize()
    dt = time.perf_counter() - t0
    print(f"sub-batches {S:2d} x {nb:5d}: {1e3 * dt / a.steps:8.2f} ms/step  {a.batch * a.nodes * a.steps / dt:.4g} node-timesteps/s", flush=True)
    for bs in hs:
        bs.close()
