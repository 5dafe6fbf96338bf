"""Duck-typed Conductor / Environment carrying exactly the attributes the
reference `step()` touches.  No reference import:
usable on the GPU box by the drop-in shim tests; `tests/golden/ref_harness.py`
adds the reference's own boundary-condition methods on top (build container only)."""
from __future__ import annotations

import types
from collections import namedtuple

import numpy as np


class _Obj:
    pass


Interface = namedtuple("Interface", ("interf_name", "comp_1", "comp_2"))
InterfaceCollection = namedtuple("InterfaceCollection", ("fluid_fluid", "fluid_solid", "solid_solid", "env_solid"))
Fluid_eq_idx = namedtuple("Fluid_eq_idx", ("velocity", "pressure", "temperature"))


def build_conductor(topo, arr, fluid_component_cls=None):
    """Duck-typed conductor + environment + qsource for the reference step()."""
    import pandas as pd
    ch_name = lambda j: f"CHAN_{j + 1}"
    sol_name = lambda l: f"SOLID_{l + 1}"

    M, S, d = topo.n_ch, topo.n_sol, topo.ndf
    E = arr.dz.shape[0]
    N = E + 1
    c = _Obj()
    c.BASE_PATH = "/nonexistent"
    c.file_input = {"EXTERNAL_FLOW": "none"}
    c.inputs = {"METHOD": topo.method, "Is_rectangular": topo.is_rectangular,
                "Height": topo.height, "Width": topo.width}
    c.grid_input = {"NELEMS": E}
    c.grid_features = {"delta_z": arr.dz.copy(), "zcoord_gauss": np.zeros(E), "N_nod": N}
    c.dict_N_equation = {"FluidComponent": 3 * M, "SolidComponent": S, "NODOFS": d,
                         "NODOFS2": 2 * d, "Total": d * N}
    c.dict_band = {"Half": 2 * d, "Main_diag": 2 * d - 1, "Full": 4 * d - 1}
    c.theta_method = topo.theta
    c.time_step = arr.dt
    c.cond_num_step = arr.num_step
    c.cond_time = [0.0, arr.dt]
    c.Delta_p_min, c.k_loc, c.lambda_v = topo.delta_p_min, topo.k_loc, topo.lambda_v
    c.dict_norm = {}
    c.EQTEIG = np.zeros(d)

    fluids, solids = [], []
    from opensc2_b200.problem import (FL_V, FL_P, FL_T, FL_RHO, FL_C0, FL_GRUN, FL_H, FL_CV, FL_F,
                         SOL_RHO, SOL_CP, SOL_K, BC_PINL, BC_POUT, BC_TINL, BC_TOUT,
                         BC_MDTIN, BC_MDTOUT, ND_VFIRST, ND_VLAST, ND_RHOFIRST, ND_RHOLAST)
    g = arr.fl_gauss
    for j in range(M):
        f = _Obj()
        f.identifier = ch_name(j)
        f.coolant = _Obj()
        f.channel = _Obj()
        f.coolant.dict_Gauss_pt = {
            "velocity": g[FL_V, j].copy(), "pressure": g[FL_P, j].copy(),
            "temperature": g[FL_T, j].copy(), "total_density": g[FL_RHO, j].copy(),
            "total_speed_of_sound": g[FL_C0, j].copy(), "Gruneisen": g[FL_GRUN, j].copy(),
            "total_enthalpy": g[FL_H, j].copy(),
            "total_isochoric_specific_heat": g[FL_CV, j].copy(),
        }
        vel = np.zeros(N)
        vel[0], vel[-1] = arr.fl_node[ND_VFIRST, j], arr.fl_node[ND_VLAST, j]
        rho = np.ones(N)
        rho[0], rho[-1] = arr.fl_node[ND_RHOFIRST, j], arr.fl_node[ND_RHOLAST, j]
        f.coolant.dict_node_pt = {"velocity": vel, "total_density": rho}
        f.coolant.operations = {
            "INTIAL": topo.ch_intial[j], "PREINL": arr.fl_bc[BC_PINL, j],
            "PREOUT": arr.fl_bc[BC_POUT, j], "TEMINL": arr.fl_bc[BC_TINL, j],
            "TEMOUT": arr.fl_bc[BC_TOUT, j], "MDTIN": arr.fl_bc[BC_MDTIN, j],
            "MDTOUT": arr.fl_bc[BC_MDTOUT, j],
        }
        f.channel.inputs = {"CROSSECTION": topo.ch_area[j], "HYDIAMETER": topo.ch_dh[j]}
        f.channel.dict_friction_factor = {False: {"total": g[FL_F, j].copy()}}
        f.channel.flow_dir = ("forward" if topo.ch_forward[j] else "backward", None)
        if fluid_component_cls is not None:   # the reference's own BC code (fluid_component.py:117-565)
            for name in ("build_th_bc_index", "impose_pressure_drop", "impose_inl_p_out_v",
                         "impose_inl_v_out_p", "_FluidComponent__impose_temperature_fw",
                         "_FluidComponent__impose_temperature_bw"):
                setattr(f, name, types.MethodType(getattr(fluid_component_cls, name), f))
            f.apply_th_bc = {1: f.impose_pressure_drop, 2: f.impose_inl_p_out_v, 3: f.impose_inl_v_out_p}
        fluids.append(f)
    for l in range(S):
        s = _Obj()
        s.identifier = sol_name(l)
        s.inputs = {"CROSSECTION": topo.sol_area[l], "COSTETA": topo.sol_costeta[l]}
        s.dict_Gauss_pt = {
            "total_density": arr.sol_gauss[SOL_RHO, l].copy(),
            "total_isobaric_specific_heat": arr.sol_gauss[SOL_CP, l].copy(),
            "total_thermal_conductivity": arr.sol_gauss[SOL_K, l].copy(),
            "Q1": arr.sol_q[0, l].copy(), "Q2": arr.sol_q[1, l].copy(),
            "temperature": np.zeros(E),
        }
        s.dict_node_pt = {}
        solids.append(s)
    for f in fluids:
        f.coolant.dict_Gauss_pt.setdefault("temperature", np.zeros(E))

    inv = lambda coll: types.SimpleNamespace(collection=coll, number=len(coll))
    c.inventory = {"FluidComponent": inv(fluids), "SolidComponent": inv(solids)}
    c.equation_index = {f.identifier: Fluid_eq_idx(j, M + j, 2 * M + j) for j, f in enumerate(fluids)}
    c.equation_index.update({s.identifier: 3 * M + l for l, s in enumerate(solids)})
    if fluid_component_cls is not None:
        for f in fluids:
            f.build_th_bc_index(c)

    env = _Obj()
    env.KIND = "Environment"
    env.inputs = {"Temperature": arr.t_env}

    ff = [Interface(f"{fluids[a].identifier}_{fluids[b].identifier}", fluids[a], fluids[b]) for a, b in topo.ff]
    fs = [Interface(f"{fluids[a].identifier}_{solids[b].identifier}", fluids[a], solids[b]) for a, b in topo.fs]
    ss = [Interface(f"{solids[a].identifier}_{solids[b].identifier}", solids[a], solids[b]) for a, b in topo.ss]
    es = [Interface(f"{env.KIND}_{solids[a].identifier}", env, solids[a]) for a in topo.es]
    c.interface = InterfaceCollection(ff, fs, ss, es)
    c.dict_interf_peri = {
        "ch_ch": {"Open": {"Gauss": {i.interf_name: arr.peri_ff[0, k].copy() for k, i in enumerate(ff)}},
                  "Close": {"Gauss": {i.interf_name: arr.peri_ff[1, k].copy() for k, i in enumerate(ff)}}},
        "ch_sol": {"Gauss": {i.interf_name: arr.peri_fs[k].copy() for k, i in enumerate(fs)}},
        "sol_sol": {"Gauss": {i.interf_name: arr.peri_ss[k].copy() for k, i in enumerate(ss)}},
        "env_sol": {"Gauss": {i.interf_name: arr.peri_es[k].copy() for k, i in enumerate(es)}},
    }
    env_htc = {}
    for k, i in enumerate(es):
        if topo.es_choice2[k] and topo.is_rectangular:
            env_htc[i.interf_name] = {"conv": {"side": arr.htc_es[k, 0].copy(),
                                               "bottom": arr.htc_es[k, 1].copy(),
                                               "top": arr.htc_es[k, 2].copy()}}
        else:
            env_htc[i.interf_name] = {"conv": arr.htc_es[k, 0].copy()}
    c.dict_Gauss_pt = {"HTC": {
        "ch_ch": {"Open": {i.interf_name: arr.htc_ff[0, k].copy() for k, i in enumerate(ff)},
                  "Close": {i.interf_name: arr.htc_ff[1, k].copy() for k, i in enumerate(ff)}},
        "ch_sol": {i.interf_name: arr.htc_fs[k].copy() for k, i in enumerate(fs)},
        "sol_sol": {"cond": {i.interf_name: arr.htc_ss[k].copy() for k, i in enumerate(ss)}},
        "env_sol": env_htc,
    }}
    choice = pd.DataFrame(0, index=[env.KIND], columns=[s.identifier for s in solids])
    for k, a in enumerate(topo.es):
        choice.at[env.KIND, solids[a].identifier] = 2 if topo.es_choice2[k] else -2
    c.dict_df_coupling = {"HTC_choice": choice}
    sysvar = np.zeros((d * N, 1))
    sysvar[:, 0] = arr.sysvar
    c.dict_Step = {"SYSVAR": sysvar, "SYSLOD": arr.syslod.copy()}
    return c, env, arr.qsource.copy()


