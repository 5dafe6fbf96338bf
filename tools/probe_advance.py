"""Small device-resident run of the full time step (K_P + K_G + K_F + K_B + K_N) for ncu captures."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from opensc2_b200.batch import BatchedStep

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--nodes", type=int, default=201)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
topo, model, tab, arr, sweep = bench.make_workload(a.batch, a.nodes, seed=1)
with BatchedStep(topo, a.batch, a.nodes) as bs:
    bs.set_model(model, [tab]); bs.upload_sweep(sweep); bs.upload_state(arr.sysvar, None); bs.upload_geometry(arr)
    bs.set_profiling(True)
    for k in range(a.steps):
        bs.advance(bench.T_START + (k + 1) * bench.DT, bench.DT, 1 + k)
    bs.synchronize()
    print({k: round(v[0] / a.steps, 4) for k, v in bs.kernel_ms().items()})
