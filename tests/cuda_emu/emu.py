"""Python driver of the CPU thread emulator (kernel LOGIC tests, CPU only)."""
import ctypes as C
import os
import subprocess

import numpy as np

from opensc2_b200._abi import Osc2Topology, c_dp, c_ip, topology_struct

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


class StepDev(C.Structure):
    """Mirror of opensc2_b200/csrc/osc2_dev.h:StepDev."""
    _fields_ = [(n, C.c_int) for n in ("B", "N", "E", "M", "S", "d", "NS", "nff", "nfs", "nss", "nes")] + [
        ("dt", C.c_double), ("t_env", C.c_double), ("num_step", C.c_int)] + [
        (n, c_dp) for n in ("dz", "fl_gauss", "fl_node", "fl_bc", "sol_gauss", "sol_q", "peri_ff", "htc_ff",
                            "peri_fs", "htc_fs", "peri_ss", "htc_ss", "peri_es", "htc_es", "qsource", "peri_ss_nodal",
                            "sysvar", "sysvar_new", "syslod", "gm", "k123", "spill",
                            "norm_solution", "norm_change", "eqteig")] + [
        ("status", c_ip), ("cap_sysmat", c_dp), ("cap_known", c_dp),
        ("norm_leaf", c_ip), ("norm_prog", c_ip), ("norm_part", c_dp), ("n_leaves", C.c_int), ("n_prog", C.c_int), ("smap", C.c_void_p), ("s_slots", C.c_int), ("cpw", C.c_int), ("NT", C.c_int),
        ("sinv", C.c_void_p), ("gm_prezeroed", C.c_int)]


_LIB = None
last_launches = 0


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "libosc2_emu.so")
        srcs = [os.path.join(HERE, f) for f in ("emu_step.cpp", "cuda_emu.h")]
        csrc = os.path.join(ROOT, "opensc2_b200", "csrc")
        srcs += [os.path.join(csrc, f) for f in sorted(os.listdir(csrc)) if f.endswith((".cuh", ".h"))]
        srcs += [os.path.join(ROOT, "include", "opensc2_b200.h")]
        exp = os.path.join(ROOT, "experiments")
        srcs += [os.path.join(exp, f) for f in ("step_kernels_split_roles.cuh", "step_kernels_fused.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            # the two kernel experiments that are not in the product build are compiled in here, so that their logic stays
            # checked (selected by OSC2_FWD_SPLIT=1 / OSC2_FUSED=1, read once per process)
            subprocess.run(["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-ffp-contract=off", "-pthread",
                            "-DOSC2_WITH_SPLIT_ROLES", "-DOSC2_WITH_FUSED_GF",
                            "-o", so, os.path.join(HERE, "emu_step.cpp")], check=True)
        _LIB = C.CDLL(so)
        _LIB.osc2_emu_gm_doubles.restype = C.c_longlong
        _LIB.osc2_emu_gm_doubles.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_longlong]
        _LIB.osc2_emu_spill_doubles.restype = C.c_longlong
        _LIB.osc2_emu_spill_doubles.argtypes = [C.c_int, C.c_int, C.c_longlong, C.c_longlong]
        assert _LIB.osc2_emu_sizeof_stepdev() == C.sizeof(StepDev)
        assert _LIB.osc2_emu_sizeof_topology() == C.sizeof(Osc2Topology)
    return _LIB


def to_dev(a):
    """[B][...] -> [...][B] contiguous (the device layout)."""
    return np.array(np.moveaxis(np.asarray(a, dtype=np.float64), 0, -1), order="C", copy=True)   # never a view: the kernels write SYSLOD in place


def from_dev(a):
    return np.ascontiguousarray(np.moveaxis(a, -1, 0))


def emu_step(topo, arr, capture=False, fast=None, split=True):
    """Run K_G, K_F, K_B, K_N on a batched StepArrays through the emulator."""
    L = lib()
    B, E = arr.dz.shape
    N, M, S, d = E + 1, topo.n_ch, topo.n_sol, topo.ndf
    total = N * d
    if fast is None:
        fast = bool(L.osc2_emu_has_fast(d))
    NS = L.osc2_emu_gm_slots(d, M, S, int(fast))
    t = topology_struct(topo)
    p = StepDev()
    p.B, p.N, p.E, p.M, p.S, p.d, p.NS = B, N, E, M, S, d, NS
    p.nff, p.nfs, p.nss, p.nes = len(topo.ff), len(topo.fs), len(topo.ss), len(topo.es)
    p.dt, p.t_env, p.num_step = arr.dt, arr.t_env, arr.num_step
    keep = {}
    for name in ("dz", "fl_gauss", "fl_node", "fl_bc", "sol_gauss", "sol_q", "peri_ff", "htc_ff",
                 "peri_fs", "htc_fs", "peri_ss", "htc_ss", "peri_es", "htc_es", "qsource", "sysvar", "syslod"):
        a = to_dev(getattr(arr, name))
        if a.size == 0:
            a = np.zeros(1)
        keep[name] = a
        setattr(p, name, a.ctypes.data_as(c_dp))
    outs = {
        "sysvar_new": np.full((total, B), np.nan),
        # fast path: zeroed once like osc2_create does (K_G skips the fields that are +0 for the whole topology)
        "gm": (np.zeros if fast else (lambda n: np.full(n, np.nan)))(L.osc2_emu_gm_doubles(d, M, S, int(fast), E, B)),
        "k123": np.zeros((3, max(p.nff, 1), E, B)), "spill": np.full(L.osc2_emu_spill_doubles(d, int(fast), N, B), np.nan),
        "norm_solution": np.zeros((d, B)), "norm_change": np.zeros((d, B)), "eqteig": np.zeros((d, B)),
    }
    for k, a in outs.items():
        setattr(p, k, a.ctypes.data_as(c_dp))
    status = np.zeros(B, dtype=np.int32)      # sticky on the device: a step only writes a failure
    p.status = status.ctypes.data_as(c_ip)
    if capture:
        outs["cap_sysmat"] = np.zeros((4 * d - 1, total, B))
        outs["cap_known"] = np.zeros((total, B))
        p.cap_sysmat = outs["cap_sysmat"].ctypes.data_as(c_dp)
        p.cap_known = outs["cap_known"].ctypes.data_as(c_dp)
    global last_launches
    last_launches = L.osc2_emu_step(C.byref(t), C.byref(p), int(fast), int(split))   # kernels launched (4 on the fused path)
    res = {
        "sysvar": from_dev(outs["sysvar_new"]), "syslod": from_dev(keep["syslod"]),
        "k123": from_dev(outs["k123"])[:, :, :p.nff], "norm_solution": from_dev(outs["norm_solution"]),
        "norm_change": from_dev(outs["norm_change"]), "eqteig": from_dev(outs["eqteig"]), "status": status,
    }
    if capture:
        res["sysmat"] = from_dev(outs["cap_sysmat"])
        res["known"] = from_dev(outs["cap_known"])
    return res
