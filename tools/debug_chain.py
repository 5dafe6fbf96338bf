import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from opensc2_b200.batch import BatchedStep
from opensc2_b200 import _lib
from opensc2_b200.model import case1_model, consistent_initial_state, synthetic_helium_table, synthetic_sweep
from opensc2_b200.problem import case1_topology, synthetic_step_arrays
from oracle import oracle as O

n, batch, dt = 41, 9, 0.1
topo = case1_topology(1); model = case1_model(False); tab = synthetic_helium_table()
arr = consistent_initial_state(topo, model, tab, n, batch, seed=11)
sweep = synthetic_sweep(topo, model, n, batch, seed=11)
host = arr.copy()
with BatchedStep(topo, batch, n) as bs:
    bs.set_model(model, [tab]); bs.upload_sweep(sweep); bs.upload_state(arr.sysvar, arr.syslod); bs.upload_geometry(arr)
    for k in range(1, 41):
        t = 9.8 + k * dt
        bs.eval_operating_conditions(t)
        ref_in = O.operating_conditions(topo, model, [tab], sweep, host.sysvar, n, t)
        bs.step(dt, k)
        try:
            bs.synchronize()
        except _lib.Osc2Error as e:
            print("step", k, "ERR", e)
        sv, sl, st = bs.download_state()
        for key, v in ref_in.items():
            setattr(host, key, v)
        host.num_step, host.dt = k, dt
        ref = O.step(topo, host)
        host.sysvar, host.syslod = ref["sysvar"], ref["syslod"]
        print(k, "status", st, "oracle", ref["status"], "maxrel sysvar", float(np.max(np.abs(sv - host.sysvar) / np.maximum(np.abs(host.sysvar), 1e-300))))
        x = host.sysvar.reshape(batch, n, 8); print("   v", x[..., :2].min(), x[..., :2].max(), "p", x[..., 2:4].min(), x[..., 2:4].max(), "T", x[..., 4:].min(), x[..., 4:].max())
