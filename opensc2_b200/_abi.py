"""ctypes mirror of include/opensc2_b200.h (structures and limits only)."""
from __future__ import annotations

import ctypes as C

import numpy as np

OSC2_MAX_CHANNELS = 16
OSC2_MAX_SOLIDS = 64
OSC2_MAX_INTERFACES = 160
OSC2_K_COUNT = 8
KERNEL_NAMES = ("transpose", "gauss", "forward", "backward", "norms", "props", "k6", "k7")

OSC2_OK = 0
OSC2_ERR_INVALID = -1
OSC2_ERR_CUDA = -2
OSC2_ERR_NO_DEVICE = -3
OSC2_ERR_SINGULAR = -4
OSC2_ERR_STATE = -5
OSC2_ERR_RANGE = -6
OSC2_ERR_TCS = -7
OSC2_STATUS_RANGE = -2000000000
OSC2_STATUS_TCS = -2000000001

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class Osc2Topology(C.Structure):
    _fields_ = [
        ("n_ch", C.c_int), ("n_sol", C.c_int),
        ("nff", C.c_int), ("nfs", C.c_int), ("nss", C.c_int), ("nes", C.c_int),
        ("ch_area", C.c_double * OSC2_MAX_CHANNELS),
        ("ch_dh", C.c_double * OSC2_MAX_CHANNELS),
        ("ch_forward", C.c_int * OSC2_MAX_CHANNELS),
        ("ch_intial", C.c_int * OSC2_MAX_CHANNELS),
        ("sol_area", C.c_double * OSC2_MAX_SOLIDS),
        ("sol_costeta", C.c_double * OSC2_MAX_SOLIDS),
        ("ff", (C.c_int * 2) * OSC2_MAX_INTERFACES),
        ("fs", (C.c_int * 2) * OSC2_MAX_INTERFACES),
        ("ss", (C.c_int * 2) * OSC2_MAX_INTERFACES),
        ("es", C.c_int * OSC2_MAX_INTERFACES),
        ("es_choice2", C.c_int * OSC2_MAX_INTERFACES),
        ("is_rectangular", C.c_int),
        ("height", C.c_double), ("width", C.c_double),
        ("delta_p_min", C.c_double), ("k_loc", C.c_double), ("lambda_v", C.c_double),
        ("theta", C.c_double),
        ("method", C.c_int),
    ]


class Osc2StepInputs(C.Structure):
    _fields_ = [(n, c_dp) for n in (
        "dz", "fl_gauss", "fl_node", "fl_bc", "sol_gauss", "sol_q", "peri_ff", "htc_ff",
        "peri_fs", "htc_fs", "peri_ss", "htc_ss", "peri_es", "htc_es", "qsource")] + [("t_env", C.c_double),
                                                                                        ("peri_ss_nodal", c_dp)]


OSC2_MAX_MATERIALS = 8
OSC2_MAX_FLUIDS = 4
OSC2_N_FLUID_PROPS = 10
_CH, _SO, _IF = OSC2_MAX_CHANNELS, OSC2_MAX_SOLIDS, OSC2_MAX_INTERFACES


class Osc2Model(C.Structure):
    """Mirror of `osc2_model` (include/opensc2_b200.h)."""
    _fields_ = [
        ("ch_fluid", C.c_int * _CH), ("ch_ifriction", C.c_int * _CH), ("ch_friction_mult", C.c_double * _CH),
        ("ch_htc_corr", C.c_int * _CH),
        ("sol_kind", C.c_int * _SO), ("sol_nmat", C.c_int * _SO),
        ("sol_mat", (C.c_int * OSC2_MAX_MATERIALS) * _SO),
        ("sol_weight", (C.c_double * OSC2_MAX_MATERIALS) * _SO), ("sol_weight_sum", C.c_double * _SO),
        ("sol_rrr", C.c_double * _SO), ("sol_tc0m", C.c_double * _SO), ("sol_bc20m", C.c_double * _SO),
        ("sol_heated", C.c_int * _SO),
        ("fs_choice", C.c_int * _IF), ("fs_htc", C.c_double * _IF), ("fs_mlt", C.c_double * _IF),
        ("ff_choice", C.c_int * _IF), ("ff_htc", C.c_double * _IF), ("ff_mlt", C.c_double * _IF),
        ("ff_thick", C.c_double * _IF),
        ("ss_choice", C.c_int * _IF), ("ss_htc", C.c_double * _IF), ("ss_mlt", C.c_double * _IF),
        ("ss_thick_rc", C.c_double * _IF), ("ss_thick_cr", C.c_double * _IF), ("ss_rcontact", C.c_double * _IF),
        ("es_htc", C.c_double * _IF),
        ("ch_roughness", C.c_double * _CH),
        ("ss_rad_den", C.c_double * _IF), ("ss_rad_mlt", C.c_double * _IF),
        ("ch_void_fraction", C.c_double * _CH),
        ("sol_c0", C.c_double * _SO),
        ("ch_lam_coef", C.c_double * _CH),
    ]


class Osc2Sweep(C.Structure):
    _fields_ = [(n, c_dp) for n in ("b_field", "eps", "q0", "q_window", "tcs", "jop")]


STEP_INPUT_FIELDS = tuple(n for n, _ in Osc2StepInputs._fields_[:-2])


def topology_struct(topo) -> Osc2Topology:
    """Flatten an `opensc2_b200.problem.Topology` into the C descriptor."""
    topo.validate()
    if topo.n_ch > OSC2_MAX_CHANNELS or topo.n_sol > OSC2_MAX_SOLIDS:
        raise ValueError("too many components for the compiled limits")
    for name in ("ff", "fs", "ss", "es"):
        if len(getattr(topo, name)) > OSC2_MAX_INTERFACES:
            raise ValueError(f"too many {name} interfaces for the compiled limits")
    t = Osc2Topology()
    t.n_ch, t.n_sol = topo.n_ch, topo.n_sol
    t.nff, t.nfs, t.nss, t.nes = len(topo.ff), len(topo.fs), len(topo.ss), len(topo.es)
    for j in range(topo.n_ch):
        t.ch_area[j] = topo.ch_area[j]
        t.ch_dh[j] = topo.ch_dh[j]
        t.ch_forward[j] = int(topo.ch_forward[j])
        t.ch_intial[j] = int(topo.ch_intial[j])
    for l in range(topo.n_sol):
        t.sol_area[l] = topo.sol_area[l]
        t.sol_costeta[l] = topo.sol_costeta[l]
    for name in ("ff", "fs", "ss"):
        dst = getattr(t, name)
        for k, (a, b) in enumerate(getattr(topo, name)):
            dst[k][0], dst[k][1] = int(a), int(b)
    for k, a in enumerate(topo.es):
        t.es[k] = int(a)
        t.es_choice2[k] = int(topo.es_choice2[k])
    t.is_rectangular = int(topo.is_rectangular)
    t.height, t.width = topo.height, topo.width
    t.delta_p_min, t.k_loc, t.lambda_v = topo.delta_p_min, topo.k_loc, topo.lambda_v
    t.theta = topo.theta
    t.method = {"BE": 0, "CN": 1}[topo.method]
    return t


def as_f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def dptr(a: np.ndarray):
    return a.ctypes.data_as(c_dp)
