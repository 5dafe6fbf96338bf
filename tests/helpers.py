"""Shared helpers for the parity tests."""
import dataclasses
import os

import numpy as np

from opensc2_b200.problem import StepArrays

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    """(topology, StepArrays inputs, dict of reference outputs) for a fixture."""
    from cases import STEP_CASES

    fac, fkw, _n, _skw = STEP_CASES[name]
    topo = fac(**fkw)
    z = np.load(os.path.join(GOLDEN_DIR, f"step_{name}.npz"))
    kw = {}
    for f in dataclasses.fields(StepArrays):
        if "in_" + f.name not in z.files:          # optional fields added after the fixture was written
            continue
        v = z["in_" + f.name]
        kw[f.name] = v if v.ndim else v.item()
    outs = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    return topo, StepArrays(**kw), outs


def assert_bits(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        rel = np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
        raise AssertionError(f"{what}: {len(bad)} of {a.size} values differ, max rel {rel:.3e}, first at {bad[0]}")


def singular_case():
    """Two pure-solid conductors (d = 1); the second has no mass term, so its
    Neumann diffusion matrix is singular: the reference raises at line 8."""
    from opensc2_b200.problem import Topology, synthetic_step_arrays

    topo = Topology(n_ch=0, n_sol=1, ch_area=(), ch_dh=(), ch_forward=(), ch_intial=(),
                    sol_area=(1e-4,), sol_costeta=(1.0,))
    arr = synthetic_step_arrays(topo, 9, batch=2, seed=2, heated=False)
    arr.sol_gauss[1, 0] = 0.0
    arr.sol_gauss[1, 2] = 100.0
    return topo, arr


def case2_with_isolated_massless_solid(n, batch, bad):
    """CASE_2 conductors; in conductor `bad` solid 5 has no mass term and no thermal contact, so its
    (uniform-mesh, Neumann) diffusion rows [1 -1], [-1/2 1 -1/2], ..., [-1 1] eliminate to an exact
    zero pivot in the last node -- the reference raises `Matrix is singular` there (tsf:763-786)."""
    from opensc2_b200.problem import case2_topology, synthetic_step_arrays

    topo = case2_topology()
    arr = synthetic_step_arrays(topo, n, batch=batch, seed=9, num_step=2, exact_squares=False, heated=False)
    l = 5
    arr.dz[bad, :] = 0.02
    arr.sol_gauss[bad, 0, l] = 0.0            # density -> mass term 0
    arr.sol_gauss[bad, 2, l] = 128.0          # conductivity
    for k, (_j, s) in enumerate(topo.fs):
        if s == l:
            arr.htc_fs[bad, k] = 0.0
    for k, (a, c) in enumerate(topo.ss):
        if l in (a, c):
            arr.htc_ss[bad, k] = 0.0
    return topo, arr, (n - 1) * topo.ndf + 3 * topo.n_ch + l


STACK_TAPE = (("ag", 2.0e-6), ("cu", 40.0e-6), ("hc276", 50.0e-6), ("re123", 1.0e-6), ("sn60pb40", 10.0e-6))


def stack_case(n_nodes=5):
    """CASE_1 topology whose first solid is a StackComponent (HTS tape: Ag / Cu / Hastelloy / REBCO / solder,
    tests/golden/make_golden_stack.py): one conductor per golden (T, B) pair, solid temperature and field
    uniform along the conductor so that every Gauss point sees exactly that pair."""
    import dataclasses

    from opensc2_b200 import model as Mdl
    from opensc2_b200.problem import case1_topology, synthetic_step_arrays

    g = np.load(os.path.join(GOLDEN_DIR, "props_stack.npz"))
    mats = {"ag": Mdl.MAT_AG, "cu": Mdl.MAT_CU_NIST, "hc276": Mdl.MAT_HC276, "re123": Mdl.MAT_RE123, "sn60pb40": Mdl.MAT_SN60PB40}
    thick = tuple(s for _k, s in STACK_TAPE)
    stack = Mdl.SolidModel(Mdl.SOLID_STACK, tuple(mats[k] for k, _s in STACK_TAPE), thick, float(np.array(thick).sum()), rrr=100.0)
    base = Mdl.case1_model(False)
    model = dataclasses.replace(base, solids=(stack,) + base.solids[1:])
    topo = case1_topology(1)
    batch = g["stack_t"].size
    arr = synthetic_step_arrays(topo, n_nodes, batch=batch, seed=4)
    sv = arr.sysvar.reshape(batch, n_nodes, topo.ndf)
    sv[:, :, 3 * topo.n_ch + 0] = g["stack_t"][:, None]
    sweep = Mdl.synthetic_sweep(topo, model, n_nodes, batch, seed=4, with_tcs=False)
    bf = np.array(sweep.b_field, copy=True)
    bf[:, 0, :] = g["stack_b"][:, None]
    sweep = dataclasses.replace(sweep, b_field=bf)
    return topo, model, Mdl.synthetic_helium_table(), arr, sweep, g
