"""Throughput probe of the generic (d = 64) path: CASE_2-sized conductors, device-resident full steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from opensc2_b200.batch import BatchedStep
from opensc2_b200.model import case2_model, consistent_initial_state, synthetic_helium_table, synthetic_sweep
from opensc2_b200.problem import case2_topology

batch, n, dt, steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 151, 0.05, 3
topo = case2_topology(); model = case2_model(topo); tab = synthetic_helium_table()
small = consistent_initial_state(topo, model, tab, n, 16, seed=1, dt=dt, length=3.0)
sw = synthetic_sweep(topo, model, n, 16, seed=1, length=3.0, with_tcs=False)
import dataclasses
arr = type(small)(**{f.name: (bench.tile(getattr(small, f.name), batch) if isinstance(getattr(small, f.name), np.ndarray) else getattr(small, f.name)) for f in dataclasses.fields(small)})
sweep = type(sw)(**{f.name: (bench.tile(getattr(sw, f.name), batch) if getattr(sw, f.name) is not None else None) for f in dataclasses.fields(sw)})
with BatchedStep(topo, batch, n) as bs:
    bs.set_model(model, [tab]); bs.upload_sweep(sweep); bs.upload_state(arr.sysvar, None); bs.upload_geometry(arr)
    for k in range(2): bs.advance(11.9 + k * dt, dt, 1 + k)
    bs.synchronize(); bs.set_profiling(True); bs.kernel_ms(reset=True); bs.timer_start()
    for k in range(steps): bs.advance(12.0 + k * dt, dt, 3 + k)
    ms = bs.timer_stop(); km = bs.kernel_ms(); bs.synchronize()
    print(f"CASE_2 d=64: {batch} x {n} nodes, {ms/steps:.2f} ms/step, {batch*n*steps/(ms*1e-3):.4g} node-timesteps/s",
          {k: round(v[0]/steps, 2) for k, v in km.items()})
