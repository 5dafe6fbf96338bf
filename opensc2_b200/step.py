"""Drop-in for the reference's `step(conductor, environment, qsource, num_step)`.

Reference: /root/reference/source_code/utility_functions/transient_solution_functions.py:186-599,
called from `Simulation.conductor_solution` (/root/reference/source_code/simulation.py:407-412).
Same signature, same attributes read, same attributes written;
the numerical work (Gauss matrices, assembly, boundary rows, row scaling,
no-pivot banded LU, substitution, norms) runs in the sm_100a kernels behind
the C ABI.  There is no CPU fallback: without the CUDA library or a GPU the
call raises.

`step()` advances one conductor (batch of 1); `step_many()` advances a list
of conductors that share a topology in ONE batched launch sequence -- that is
the call a parameter sweep should make.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

from .batch import BatchedStep
from .problem import (BC_MDTIN, BC_MDTOUT, BC_PINL, BC_POUT, BC_TINL, BC_TOUT, FL_C0, FL_CV, FL_F,
                      FL_GRUN, FL_H, FL_P, FL_RHO, FL_T, FL_V, N_BC, N_FL_GAUSS, N_ND, N_SOL_GAUSS,
                      ND_RHOFIRST, ND_RHOLAST, ND_VFIRST, ND_VLAST, SOL_CP, SOL_K, SOL_RHO, StepArrays,
                      Topology)

_METHOD_THETA = {"BE": 1.0, "CN": 0.5}
_HANDLES: Dict[Tuple, BatchedStep] = {}


def topology_of(conductor) -> Topology:
    """Flatten what step() reads from the Conductor object graph."""
    fluids = conductor.inventory["FluidComponent"].collection
    solids = conductor.inventory["SolidComponent"].collection
    fidx = {id(f): j for j, f in enumerate(fluids)}
    sidx = {id(s): l for l, s in enumerate(solids)}
    # the kernels rely on the reference's equation order (conductor.py:897-946)
    M = len(fluids)
    for j, f in enumerate(fluids):
        e = conductor.equation_index[f.identifier]
        if (e.velocity, e.pressure, e.temperature) != (j, M + j, 2 * M + j):
            raise ValueError(f"unexpected equation index for {f.identifier}: {e}")
    for l, s in enumerate(solids):
        if conductor.equation_index[s.identifier] != 3 * M + l:
            raise ValueError(f"unexpected equation index for {s.identifier}")
    method = conductor.inputs["METHOD"]
    if method not in ("BE", "CN"):
        raise ValueError(f"METHOD {method!r}: only BE and CN are implemented on the device (AM4 is not)")
    itf = conductor.interface
    choice = conductor.dict_df_coupling["HTC_choice"]
    es = tuple(sidx[id(i.comp_2)] for i in itf.env_solid)
    es_choice2 = tuple(bool(choice.at[i.comp_1.KIND, i.comp_2.identifier] == 2) for i in itf.env_solid)
    return Topology(
        n_ch=M, n_sol=len(solids),
        ch_area=tuple(float(f.channel.inputs["CROSSECTION"]) for f in fluids),
        ch_dh=tuple(float(f.channel.inputs["HYDIAMETER"]) for f in fluids),
        ch_forward=tuple(f.channel.flow_dir[0] == "forward" for f in fluids),
        ch_intial=tuple(_positive_intial(f) for f in fluids),
        sol_area=tuple(float(s.inputs["CROSSECTION"]) for s in solids),
        sol_costeta=tuple(float(s.inputs["COSTETA"]) for s in solids),
        ff=tuple((fidx[id(i.comp_1)], fidx[id(i.comp_2)]) for i in itf.fluid_fluid),
        fs=tuple((fidx[id(i.comp_1)], sidx[id(i.comp_2)]) for i in itf.fluid_solid),
        ss=tuple((sidx[id(i.comp_1)], sidx[id(i.comp_2)]) for i in itf.solid_solid),
        es=es, es_choice2=es_choice2,
        is_rectangular=bool(conductor.inputs["Is_rectangular"]),
        height=float(conductor.inputs["Height"] or 0.0), width=float(conductor.inputs["Width"] or 0.0),
        delta_p_min=float(conductor.Delta_p_min), k_loc=float(conductor.k_loc),
        lambda_v=float(conductor.lambda_v), theta=float(conductor.theta_method), method=method,
    )


def _positive_intial(f) -> int:
    """INTIAL in {1, 2, 3}.  The negative flags make the reference read time-interpolated boundary values from the
    EXTERNAL_FLOW workbook inside impose_pressure_drop / impose_inl_p_out_v / impose_inl_v_out_p
    (fluid_component.py:255-565, `get_from_xlsx` branches) -- not implemented here, and not to be confused with
    their positive twins, so they are refused."""
    flag = int(f.coolant.operations["INTIAL"])
    if flag not in (1, 2, 3):
        raise ValueError(f"{f.identifier}: INTIAL = {flag}: only 1, 2, 3 are implemented (negative flags take their "
                         "boundary values from the EXTERNAL_FLOW file at every step)")
    return flag


# boundary values each hydraulic mode reads (fluid_component.py:279-287, 383-391, 487-495 + the temperature rows)
_BC_KEYS_READ = {1: ("PREINL", "PREOUT", "TEMINL", "TEMOUT"), 2: ("PREINL", "MDTOUT", "TEMINL", "TEMOUT"),
                 3: ("MDTIN", "PREOUT", "TEMINL", "TEMOUT")}


def pack(conductor, environment, qsource, num_step: int, topo: Topology) -> StepArrays:
    """One conductor's step() inputs as flat arrays (no batch axis)."""
    fluids = conductor.inventory["FluidComponent"].collection
    solids = conductor.inventory["SolidComponent"].collection
    M, S = topo.n_ch, topo.n_sol
    dz = np.asarray(conductor.grid_features["delta_z"], dtype=np.float64)
    E = dz.shape[0]
    N = E + 1
    fl_gauss = np.empty((N_FL_GAUSS, M, E))
    fl_node = np.empty((N_ND, M))
    fl_bc = np.empty((N_BC, M))
    keys = ((FL_V, "velocity"), (FL_P, "pressure"), (FL_T, "temperature"), (FL_RHO, "total_density"),
            (FL_C0, "total_speed_of_sound"), (FL_GRUN, "Gruneisen"), (FL_H, "total_enthalpy"),
            (FL_CV, "total_isochoric_specific_heat"))
    for j, f in enumerate(fluids):
        g = f.coolant.dict_Gauss_pt
        for q, key in keys:
            fl_gauss[q, j] = g[key]
        fl_gauss[FL_F, j] = f.channel.dict_friction_factor[False]["total"]
        nd = f.coolant.dict_node_pt
        fl_node[ND_VFIRST, j], fl_node[ND_VLAST, j] = nd["velocity"][0], nd["velocity"][-1]
        fl_node[ND_RHOFIRST, j], fl_node[ND_RHOLAST, j] = nd["total_density"][0], nd["total_density"][-1]
        op = f.coolant.operations
        for q, key in ((BC_PINL, "PREINL"), (BC_POUT, "PREOUT"), (BC_TINL, "TEMINL"), (BC_TOUT, "TEMOUT"),
                       (BC_MDTIN, "MDTIN"), (BC_MDTOUT, "MDTOUT")):
            if key in _BC_KEYS_READ[topo.ch_intial[j]]:
                fl_bc[q, j] = op[key]                       # a missing key fails loudly (KeyError), like the reference
            else:
                fl_bc[q, j] = op[key] if key in op else 0.0   # never read in this mode
    sol_gauss = np.empty((N_SOL_GAUSS, S, E))
    sol_q = np.zeros((2, S, E, 2))
    for l, s in enumerate(solids):
        g = s.dict_Gauss_pt
        sol_gauss[SOL_RHO, l] = g["total_density"]
        sol_gauss[SOL_CP, l] = g["total_isobaric_specific_heat"]
        sol_gauss[SOL_K, l] = g["total_thermal_conductivity"]
        for q, key in enumerate(("Q1", "Q2")):
            a = np.asarray(g[key], dtype=np.float64)
            if a.ndim == 1:          # only the present column exists after the first step
                sol_q[q, l, :, 0] = a
            else:
                sol_q[q, l, :, :a.shape[1]] = a[:, :2]
    itf = conductor.interface
    peri, htc = conductor.dict_interf_peri, conductor.dict_Gauss_pt["HTC"]
    nff, nfs, nss, nes = len(itf.fluid_fluid), len(itf.fluid_solid), len(itf.solid_solid), len(itf.env_solid)
    peri_ff, htc_ff = np.empty((2, nff, E)), np.empty((2, nff, E))
    for k, i in enumerate(itf.fluid_fluid):
        for q, oc in enumerate(("Open", "Close")):
            peri_ff[q, k] = peri["ch_ch"][oc]["Gauss"][i.interf_name]
            htc_ff[q, k] = htc["ch_ch"][oc][i.interf_name]
    peri_fs, htc_fs = np.empty((nfs, E)), np.empty((nfs, E))
    for k, i in enumerate(itf.fluid_solid):
        peri_fs[k] = peri["ch_sol"]["Gauss"][i.interf_name]
        htc_fs[k] = htc["ch_sol"][i.interf_name]
    peri_ss, htc_ss = np.empty((nss, E)), np.empty((nss, E))
    for k, i in enumerate(itf.solid_solid):
        peri_ss[k] = peri["sol_sol"]["Gauss"][i.interf_name]
        htc_ss[k] = htc["sol_sol"]["cond"][i.interf_name]
    peri_es, htc_es = np.empty((nes, E)), np.zeros((nes, 3, E))
    for k, i in enumerate(itf.env_solid):
        peri_es[k] = peri["env_sol"]["Gauss"][i.interf_name]
        h = htc["env_sol"][i.interf_name]["conv"]
        if topo.es_choice2[k] and topo.is_rectangular:
            htc_es[k, 0], htc_es[k, 1], htc_es[k, 2] = h["side"], h["bottom"], h["top"]
        else:
            htc_es[k, 0] = h
    qs = np.asarray(qsource, dtype=np.float64)
    if qs.shape != (N, S):
        raise ValueError(f"qsource: expected shape {(N, S)}, got {qs.shape}")
    return StepArrays(
        dz=dz, fl_gauss=fl_gauss, fl_node=fl_node, fl_bc=fl_bc, sol_gauss=sol_gauss, sol_q=sol_q,
        peri_ff=peri_ff, htc_ff=htc_ff, peri_fs=peri_fs, htc_fs=htc_fs, peri_ss=peri_ss, htc_ss=htc_ss,
        peri_es=peri_es, htc_es=htc_es, qsource=qs,
        sysvar=np.array(conductor.dict_Step["SYSVAR"][:, 0], dtype=np.float64),
        syslod=np.array(conductor.dict_Step["SYSLOD"], dtype=np.float64),
        t_env=float(environment.inputs["Temperature"]), dt=float(conductor.time_step), num_step=int(num_step),
    )


def unpack(conductor, out: Dict[str, np.ndarray], b: int, topo: Topology) -> None:
    """Write what the reference step() writes (tsf:564-599, reorganize_th_solution tsf:973-1030)."""
    fluids = conductor.inventory["FluidComponent"].collection
    solids = conductor.inventory["SolidComponent"].collection
    M, d = topo.n_ch, topo.ndf
    old_t = {f.identifier: f.coolant.dict_Gauss_pt["temperature"] for f in fluids}
    old_t.update({s.identifier: s.dict_Gauss_pt["temperature"] for s in solids})
    conductor.dict_Step["SYSVAR"][:, 0] = out["sysvar"][b]
    conductor.dict_Step["SYSLOD"][...] = out["syslod"][b]
    for q, key in enumerate(("K1", "K2", "K3")):
        dst = conductor.dict_Gauss_pt.setdefault(key, {})
        for k, i in enumerate(conductor.interface.fluid_fluid):
            dst[i.interf_name] = out["k123"][b, q, k].copy()
    conductor.dict_norm["Solution"] = out["norm_solution"][b].copy()
    conductor.dict_norm["Change"] = out["norm_change"][b].copy()
    conductor.EQTEIG = out["eqteig"][b].copy()
    sysvar = conductor.dict_Step["SYSVAR"]
    for j, f in enumerate(fluids):
        nd = f.coolant.dict_node_pt
        nd["velocity"] = sysvar[j::d, 0].copy()
        nd["pressure"] = sysvar[M + j::d, 0].copy()
        nd["temperature"] = sysvar[2 * M + j::d, 0].copy()
        f.coolant.dict_Gauss_pt["temperature_change"] = (
            nd["temperature"][:-1] + nd["temperature"][1:]) / 2. - old_t[f.identifier]
    for l, s in enumerate(solids):
        s.dict_node_pt["temperature"] = sysvar[3 * M + l::d, 0].copy()
        s.dict_Gauss_pt["temperature_change"] = (
            s.dict_node_pt["temperature"][:-1] + s.dict_node_pt["temperature"][1:]) / 2. - old_t[s.identifier]


def _handle(topo: Topology, batch: int, n_nodes: int, device: int) -> BatchedStep:
    key = (topo, batch, n_nodes, device)
    h = _HANDLES.get(key)
    if h is None:
        if len(_HANDLES) >= 8:           # bounded cache of device contexts
            _HANDLES.pop(next(iter(_HANDLES))).close()
        h = _HANDLES[key] = BatchedStep(topo, batch, n_nodes, device=device)
    return h


def release() -> None:
    """Free the cached device contexts."""
    while _HANDLES:
        _HANDLES.popitem()[1].close()


def step_many(conductors: Sequence, environment, qsources: Sequence, num_step: int, device: int = 0) -> None:
    """`step()` for a list of conductors sharing a topology and a mesh size:
    one batched device call; each conductor is mutated exactly as the reference does."""
    conductors = list(conductors)
    if not conductors:
        return
    topo = topology_of(conductors[0])
    arrs: List[StepArrays] = []
    for c, qs in zip(conductors, qsources):
        if topology_of(c) != topo:
            raise ValueError("step_many: conductors must share topology, areas and time scheme")
        if int(c.cond_num_step) != int(num_step):
            raise ValueError("num_step differs from conductor.cond_num_step")
        arrs.append(pack(c, environment, qs, num_step, topo))
    dt = arrs[0].dt
    if any(a.dt != dt for a in arrs):
        raise ValueError("step_many: conductors must share the time step")
    n_nodes = arrs[0].n_nodes
    if any(a.n_nodes != n_nodes for a in arrs):
        raise ValueError("step_many: conductors must share the number of nodes")
    kw = {}
    for name in StepArrays.__dataclass_fields__:
        v0 = getattr(arrs[0], name)
        kw[name] = np.stack([getattr(a, name) for a in arrs]) if isinstance(v0, np.ndarray) else v0
    batch = StepArrays(**kw)
    h = _handle(topo, len(arrs), n_nodes, device)
    out = h.run(batch)       # raises SingularMatrixError (a ValueError) like gredub/gbacsb
    for b, c in enumerate(conductors):
        unpack(c, out, b, topo)


def step(conductor, environment, qsource, num_step):
    """Reference-compatible single-conductor entry point (returns None, mutates `conductor`)."""
    step_many([conductor], environment, [qsource], num_step)
