"""tests/golden/time_step.npz: the UNMODIFIED reference get_time_step (tsf:38-184) on random inputs.
Build-container only:  python tests/golden/make_golden_time_step.py"""
import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402


def main():
    T = H._import_reference()
    rng = np.random.default_rng(5)
    rows = []
    for _ in range(400):
        mode = int(rng.choice([0, 1, -1, 2, 3]))
        nf, ns = int(rng.integers(1, 4)), int(rng.integers(0, 4))
        d = 3 * nf + ns
        eq = 10.0 ** rng.uniform(-3, 3, d)
        prev = float(10.0 ** rng.uniform(-3, 0))
        t = float(rng.uniform(0.0, 120.0))
        num_step = int(rng.choice([1, 2, 7]))
        stpmin, tend, eigtim, shift = float(10.0 ** rng.uniform(-3, -1)), 100.0, float(rng.uniform(0.5, 2.0)), float(rng.uniform(0, 20))
        c = types.SimpleNamespace(time_step=prev, EQTEIG=eq, EIGTIM=eigtim, cond_time=[0.0, t],
                                  dict_N_equation={"NODOFS": d, "FluidComponent": 3 * nf},
                                  inventory={"FluidComponent": types.SimpleNamespace(number=nf),
                                             "SolidComponent": types.SimpleNamespace(number=ns)})
        ti = {"STPMIN": stpmin, "TEND": tend, "IADAPTIME": mode, "TIME_SHIFT": shift}
        with contextlib.redirect_stdout(io.StringIO()):
            T.get_time_step(c, ti, num_step)
        rows.append([mode, nf, ns, prev, t, num_step, stpmin, tend, eigtim, shift, c.time_step] + list(eq) + [0.0] * (13 - d))
    np.savez_compressed(os.path.join(HERE, "time_step.npz"), rows=np.array(rows))
    print("wrote", len(rows))


if __name__ == "__main__":
    main()
