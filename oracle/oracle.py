"""ctypes wrapper of the CPU ORACLE (test infrastructure, not product code).

Only tests/, `__graft_entry__.smoke()` and bench.py's cpu_baseline /
`--impl reference` legs import this module.  See oracle/step_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class _Topo(C.Structure):
    _fields_ = [
        ("n_ch", C.c_int), ("n_sol", C.c_int),
        ("nff", C.c_int), ("nfs", C.c_int), ("nss", C.c_int), ("nes", C.c_int),
        ("ch_area", c_dp), ("ch_dh", c_dp), ("ch_forward", c_ip), ("ch_intial", c_ip),
        ("sol_area", c_dp), ("sol_costeta", c_dp),
        ("ff", c_ip), ("fs", c_ip), ("ss", c_ip), ("es", c_ip), ("es_choice2", c_ip),
        ("is_rectangular", C.c_int),
        ("height", C.c_double), ("width", C.c_double),
        ("delta_p_min", C.c_double), ("k_loc", C.c_double), ("lambda_v", C.c_double),
        ("theta", C.c_double), ("method", C.c_int),
    ]


class _StepIO(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int),
        ("dz", c_dp), ("fl_gauss", c_dp), ("fl_node", c_dp), ("fl_bc", c_dp),
        ("sol_gauss", c_dp), ("sol_q", c_dp), ("peri_ff", c_dp), ("htc_ff", c_dp),
        ("peri_fs", c_dp), ("htc_fs", c_dp), ("peri_ss", c_dp), ("htc_ss", c_dp),
        ("peri_es", c_dp), ("htc_es", c_dp), ("qsource", c_dp),
        ("t_env", C.c_double), ("dt", C.c_double), ("num_step", C.c_int),
        ("sysvar", c_dp), ("syslod", c_dp),
        ("k123", c_dp), ("norm_solution", c_dp), ("norm_change", c_dp), ("eqteig", c_dp),
        ("sysmat", c_dp), ("known", c_dp),
        ("status", C.c_int),
    ]


def build(force: bool = False) -> str:
    """Compile oracle/libosc2_oracle.so with gcc (Makefile beside this file)."""
    so = os.path.join(_HERE, "libosc2_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("step_oracle.c", "step_oracle.h", "props_oracle.c", "props_oracle.h")]
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.osc2o_step.argtypes = [C.POINTER(_Topo), C.POINTER(_StepIO)]
        _LIB.osc2o_step.restype = C.c_int
        _LIB.osc2o_step_batch.argtypes = [C.POINTER(_Topo), C.POINTER(_StepIO), C.c_int, C.c_int, c_ip]
        _LIB.osc2o_step_batch.restype = C.c_int
        _LIB.osc2o_pairwise_sum.argtypes = [c_dp, C.c_long, C.c_long]
        _LIB.osc2o_pairwise_sum.restype = C.c_double
        _LIB.osc2o_fit.argtypes = [C.c_int, C.c_long, c_dp, c_dp, c_dp, C.c_double, C.c_double, c_dp]
        _LIB.osc2o_fit.restype = C.c_int
    return _LIB


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_dp)


def _i(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.int32).reshape(-1))
    if a.size == 0:
        a = np.zeros(1, dtype=np.int32)
    return a, a.ctypes.data_as(c_ip)


def _topo_struct(topo):
    keep = []
    t = _Topo()
    t.n_ch, t.n_sol = topo.n_ch, topo.n_sol
    t.nff, t.nfs, t.nss, t.nes = len(topo.ff), len(topo.fs), len(topo.ss), len(topo.es)
    for name, val, conv in (
        ("ch_area", topo.ch_area, _d), ("ch_dh", topo.ch_dh, _d),
        ("ch_forward", [int(x) for x in topo.ch_forward], _i), ("ch_intial", topo.ch_intial, _i),
        ("sol_area", topo.sol_area if topo.n_sol else [0.0], _d),
        ("sol_costeta", topo.sol_costeta if topo.n_sol else [1.0], _d),
        ("ff", topo.ff, _i), ("fs", topo.fs, _i), ("ss", topo.ss, _i), ("es", topo.es, _i),
        ("es_choice2", [int(x) for x in topo.es_choice2], _i),
    ):
        arr, ptr = conv(val)
        keep.append(arr)
        setattr(t, name, ptr)
    t.is_rectangular = int(topo.is_rectangular)
    t.height, t.width = topo.height, topo.width
    t.delta_p_min, t.k_loc, t.lambda_v = topo.delta_p_min, topo.k_loc, topo.lambda_v
    t.theta = topo.theta
    t.method = {"BE": 0, "CN": 1}[topo.method]
    return t, keep


_IN = ("dz", "fl_gauss", "fl_node", "fl_bc", "sol_gauss", "sol_q", "peri_ff", "htc_ff",
       "peri_fs", "htc_fs", "peri_ss", "htc_ss", "peri_es", "htc_es", "qsource")


def step(topo, arr, want_system: bool = False, n_threads: int = 0):
    """Reference-equivalent step() on CPU.  `arr` may be batched (leading axis).

    Returns a dict: sysvar, syslod (new), k123, norm_solution, norm_change,
    eqteig, status and, if `want_system`, the pre-solve sysmat/known.
    """
    L = lib()
    d = topo.ndf
    batched = arr.dz.ndim == 2
    B = arr.dz.shape[0] if batched else 1
    E = arr.dz.shape[-1]
    N = E + 1
    total = d * N
    lead = (B,) if batched else ()
    t, keep = _topo_struct(topo)
    io = _StepIO()
    io.n_nodes = N
    for name in _IN:
        a, p = _d(getattr(arr, name))
        if a.size == 0:
            a = np.zeros(1)
            p = a.ctypes.data_as(c_dp)
        keep.append(a)
        setattr(io, name, p)
    io.t_env, io.dt, io.num_step = float(arr.t_env), float(arr.dt), int(arr.num_step)
    out = {
        "sysvar": np.array(arr.sysvar, dtype=np.float64, order="C", copy=True),
        "syslod": np.array(arr.syslod, dtype=np.float64, order="C", copy=True),
        "k123": np.zeros(lead + (3, max(len(topo.ff), 0), E)),
        "norm_solution": np.zeros(lead + (d,)),
        "norm_change": np.zeros(lead + (d,)),
        "eqteig": np.zeros(lead + (d,)),
    }
    if want_system:
        out["sysmat"] = np.zeros(lead + (4 * d - 1, total))
        out["known"] = np.zeros(lead + (total,))
    for name, a in out.items():
        if a.size:
            setattr(io, name, a.ctypes.data_as(c_dp))
    status = np.zeros(B, dtype=np.int32)
    if batched:
        L.osc2o_step_batch(C.byref(t), C.byref(io), B, n_threads, status.ctypes.data_as(c_ip))
    else:
        status[0] = L.osc2o_step(C.byref(t), C.byref(io))
    out["status"] = status if batched else int(status[0])
    return out


def pairwise_sum(a, stride: int = 1) -> float:
    a = np.ascontiguousarray(a, dtype=np.float64)
    n = (a.size + stride - 1) // stride
    return lib().osc2o_pairwise_sum(a.ctypes.data_as(c_dp), n, stride)


FITS = {"k_cu": 0, "cp_cu": 1, "k_nb3sn": 2, "cp_nb3sn": 3, "tc_nb3sn": 4, "k_ss": 5, "cp_ss": 6, "k_ge": 7,
        "cp_ge": 8, "friction_124": 9, "nusselt_1": 10, "nusselt_212": 11, "rhoe_cu": 12,
        "k_al": 13, "cp_al": 14, "k_re123": 15, "cp_re123": 16, "friction_122": 17,
        "k_ag": 18, "cp_ag": 19, "k_hc276": 20, "cp_hc276": 21, "k_sn60pb40": 22, "cp_sn60pb40": 23, "friction_simple": 24, "k_nbti": 25, "cp_nbti": 26, "tc_nbti": 27, "nusselt_more": 28, "k_mgb2": 29, "cp_mgb2": 30, "density": 31,
        "friction_hole_newton": 32}


def fit(name, x, y=None, z=None, p1=0.0, p2=0.0, out_init=None):
    """Element-wise material / correlation fit of oracle/props_oracle.c (see props_oracle.h).  `out_init`: a
    fourth per-element argument for the fits that need one (cp_nbti: the critical temperature)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(x if y is None else y, dtype=np.float64)
    z = np.ascontiguousarray(x if z is None else z, dtype=np.float64)
    out = np.empty_like(x) if out_init is None else np.array(out_init, dtype=np.float64, copy=True)
    rc = lib().osc2o_fit(FITS[name], x.size, x.ctypes.data_as(c_dp), y.ctypes.data_as(c_dp), z.ctypes.data_as(c_dp),
                         float(p1), float(p2), out.ctypes.data_as(c_dp))
    assert rc == 0
    return out


# ---- operating conditions at the Gauss points (oracle/props_oracle.c) -----------------
class _FluidTab(C.Structure):
    _fields_ = [("n_p", C.c_int), ("n_t", C.c_int), ("p_min", C.c_double), ("p_step", C.c_double),
                ("t_min", C.c_double), ("t_step", C.c_double), ("props", c_dp)]


class _OpcondIO(C.Structure):
    _fields_ = [("batch", C.c_int), ("n_nodes", C.c_int), ("time", C.c_double), ("sysvar", c_dp),
                ("fluid", _FluidTab * 4)] + [(n, c_dp) for n in (
        "b_field", "eps", "q0", "q_window", "tcs", "peri_ss_nodal",
        "fl_gauss", "fl_node", "sol_gauss", "sol_q", "htc_ff", "htc_fs", "htc_ss", "htc_es")]


def operating_conditions(topo, model, tables, sweep, sysvar, n_nodes, time, n_threads=1, peri_ss_nodal=None):
    """CPU restatement of what K_P computes.  `sysvar`: (B, N*d); returns the dict of
    step-input arrays (batch leading) that `osc2_download_step_inputs` returns."""
    from opensc2_b200._abi import topology_struct      # struct definitions only (the C header's mirror)

    L = lib()
    L.osc2o_operating_conditions.restype = C.c_int
    sysvar = np.ascontiguousarray(sysvar, dtype=np.float64)
    B = sysvar.shape[0]
    E, M, S = n_nodes - 1, topo.n_ch, topo.n_sol
    t = topology_struct(topo)
    m = model.struct(topo)
    io = _OpcondIO()
    io.batch, io.n_nodes, io.time = B, n_nodes, float(time)
    io.sysvar = sysvar.ctypes.data_as(c_dp)
    keep = [sysvar]
    for i, tab in enumerate(tables):
        io.fluid[i].n_p, io.fluid[i].n_t = tab.props.shape[1], tab.props.shape[2]
        io.fluid[i].p_min, io.fluid[i].p_step = tab.p_min, tab.p_step
        io.fluid[i].t_min, io.fluid[i].t_step = tab.t_min, tab.t_step
        io.fluid[i].props = tab.props.ctypes.data_as(c_dp)
    for name in ("b_field", "eps", "q0", "q_window", "tcs"):
        a = getattr(sweep, name)
        if a is None or a.size == 0:
            setattr(io, name, None)
            continue
        a = np.ascontiguousarray(a, dtype=np.float64)
        keep.append(a)
        setattr(io, name, a.ctypes.data_as(c_dp))
    io.peri_ss_nodal = None
    if peri_ss_nodal is not None:
        pn = np.ascontiguousarray(peri_ss_nodal, dtype=np.float64)
        keep.append(pn)
        io.peri_ss_nodal = pn.ctypes.data_as(c_dp)
    out = {"fl_gauss": np.zeros((B, 9, M, E)), "fl_node": np.zeros((B, 4, M)), "sol_gauss": np.zeros((B, 3, S, E)),
           "sol_q": np.zeros((B, 2, S, E, 2)), "htc_ff": np.zeros((B, 2, len(topo.ff), E)),
           "htc_fs": np.zeros((B, len(topo.fs), E)), "htc_ss": np.zeros((B, len(topo.ss), E)),
           "htc_es": np.zeros((B, len(topo.es), 3, E))}
    dummy = np.zeros(1)
    for k, a in out.items():
        setattr(io, k, (a if a.size else dummy).ctypes.data_as(c_dp))
    rc = L.osc2o_operating_conditions(C.byref(t), C.byref(m), C.byref(io), int(n_threads))
    assert rc == 0
    return out
