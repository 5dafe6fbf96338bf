"""Smallest end-to-end case for compute-sanitizer: fast path (d = 8), generic path (d = 9) and the
experimental kernels, a few nodes, a handful of conductors."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from opensc2_b200.batch import BatchedStep
from opensc2_b200.model import (case1_model, case2_mini_model, consistent_initial_state, synthetic_helium_table,
                                synthetic_sweep)
from opensc2_b200.problem import case1_topology, case2_mini_topology

tab = synthetic_helium_table()
for topo, model, n, batch in ((case1_topology(1), case1_model(True), 9, 6), (case2_mini_topology(), case2_mini_model(), 7, 3)):
    arr = consistent_initial_state(topo, model, tab, n, batch, seed=1)
    sweep = synthetic_sweep(topo, model, n, batch, seed=1, with_tcs=topo.ndf == 8)
    with BatchedStep(topo, batch, n) as bs:
        bs.set_model(model, [tab]); bs.upload_sweep(sweep); bs.upload_state(arr.sysvar, arr.syslod); bs.upload_geometry(arr)
        for k in range(1, 3):
            bs.advance(9.9 + 0.1 * k, 0.1, k)
        bs.synchronize()
        sv, sl, st = bs.download_state()
        d = bs.download_diagnostics()
        assert np.all(st == 0) and np.all(np.isfinite(sv))
print("sanitize case ok")
