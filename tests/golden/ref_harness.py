"""Run the UNMODIFIED reference `step()` in-process on a (Topology, StepArrays).

Only usable where `/root/reference` exists (the build container).  Used by
`make_golden.py` to generate the committed fixtures and by the optional
`tests/test_oracle_vs_reference.py` live check.  Nothing under `-m gpu`,
`bench.py` or `smoke()` imports this module.

Recipe: SURVEY.md Appendix D -- stub the absent third-party modules, import
`utility_functions.transient_solution_functions`, build plain objects carrying
exactly the attributes `step` touches, bind the reference's own boundary
condition methods onto the fake channels.
"""
from __future__ import annotations

import os
import sys
import types
from collections import namedtuple

import numpy as np

REF_SRC = "/root/reference/source_code"


def reference_available() -> bool:
    return os.path.isdir(REF_SRC)


_T = None


def _import_reference():
    global _T
    if _T is not None:
        return _T
    import warnings

    warnings.simplefilter("ignore", SyntaxWarning)
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)

    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    try:
        import openpyxl  # noqa: F401
    except Exception:
        stub("openpyxl", load_workbook=lambda *a, **k: None)
    try:
        import CoolProp.CoolProp  # noqa: F401
    except Exception:
        cp = stub("CoolProp")
        cpcp = stub("CoolProp.CoolProp", PropsSI=lambda *a, **k: (_ for _ in ()).throw(RuntimeError("CoolProp absent")))
        cp.CoolProp = cpcp
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = stub("matplotlib", use=lambda *a, **k: None)
        plt = stub("matplotlib.pyplot")
        mpl.pyplot = plt
    try:
        import line_profiler  # noqa: F401
    except Exception:
        stub("line_profiler", LineProfiler=object)

    import utility_functions.transient_solution_functions as T

    _T = T
    return T


def build_conductor(topo, arr):
    """Duck-typed conductor + environment + qsource for the reference step()."""
    _import_reference()
    from fluid_component import FluidComponent

    root_tests = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root_tests not in sys.path:
        sys.path.insert(0, root_tests)
    import fake_conductor

    return fake_conductor.build_conductor(topo, arr, fluid_component_cls=FluidComponent)


def run_reference_step(topo, arr):
    """Returns dict with SYSVAR, SYSLOD, pre-solve SYSMAT/Known, norms, EQTEIG, K1..K3."""
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if root not in sys.path:
        sys.path.insert(0, root)
    T = _import_reference()
    cond, env, qsource = build_conductor(topo, arr)
    cap = {}
    g0, b0 = T.gredub, T.gbacsb

    def gredub(conductor, A):
        cap["SYSMAT"] = A.copy()
        return g0(conductor, A)

    def gbacsb(conductor, A, B):
        cap["Known"] = B.copy()
        return b0(conductor, A, B)

    T.gredub, T.gbacsb = gredub, gbacsb
    try:
        T.step(cond, env, qsource, arr.num_step)
    finally:
        T.gredub, T.gbacsb = g0, b0
    out = {
        "sysvar": cond.dict_Step["SYSVAR"][:, 0].copy(),
        "syslod": cond.dict_Step["SYSLOD"].copy(),
        "sysmat": cap["SYSMAT"], "known": cap["Known"],
        "norm_solution": np.asarray(cond.dict_norm["Solution"]).copy(),
        "norm_change": np.asarray(cond.dict_norm["Change"]).copy(),
        "eqteig": np.asarray(cond.EQTEIG).copy(),
    }
    nff = len(topo.ff)
    E = arr.dz.shape[0]
    k123 = np.zeros((3, nff, E))
    for k, i in enumerate(cond.interface.fluid_fluid):
        for q, key in enumerate(("K1", "K2", "K3")):
            k123[q, k] = cond.dict_Gauss_pt[key][i.interf_name]
    out["k123"] = k123
    return out
