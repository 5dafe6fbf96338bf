"""tests/golden/save_properties/*.tsv: files written by the UNMODIFIED reference `save_properties`
(utility_functions/output.py:8-92) for a small CASE_1-shaped conductor whose nodal coolant properties
were produced by the reference `Coolant` methods (PropsSI := the table lookup of this build).
Build-container only:  python tests/golden/make_golden_output.py"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402
from make_golden_opcond import ALIAS_TO_PROP  # noqa: E402
from opensc2_b200.model import case1_model, consistent_initial_state, synthetic_helium_table, synthetic_sweep, table_lookup  # noqa: E402
from opensc2_b200.problem import case1_topology  # noqa: E402


def main():
    H._import_reference()
    import coolant as coolant_mod
    from channel import Channel
    from coolant import Coolant
    from utility_functions.output import save_properties

    topo, model, tab = case1_topology(1), case1_model(True), synthetic_helium_table()
    n = 7
    arr = consistent_initial_state(topo, model, tab, n, 1, seed=12)
    sweep = synthetic_sweep(topo, model, n, 1, seed=12)
    rng = np.random.default_rng(1)
    x = arr.sysvar[0].reshape(n, topo.ndf)
    x += rng.uniform(0.0, 1e-3, x.shape) * np.abs(x)
    coolant_mod.PropsSI = lambda alias, _t, temperature, _p, pressure, _f: table_lookup(tab, ALIAS_TO_PROP[alias], pressure, temperature)
    aliases = dict(isobaric_expansion_coefficient="isobaric_expansion_coefficient",
                   isothermal_compressibility="isothermal_compressibility", Prandtl="Prandtl",
                   total_dynamic_viscosity="viscosity", total_enthalpy="Hmass", total_isobaric_specific_heat="Cpmass",
                   total_isochoric_specific_heat="Cvmass", total_speed_of_sound="speed_of_sound",
                   total_thermal_conductivity="conductivity")
    M = topo.n_ch
    zcoord = np.linspace(0.0, 10.0, n)
    cond = types.SimpleNamespace(grid_features={"N_nod": n, "zcoord": zcoord})
    fluids = []
    for j in range(M):
        co = object.__new__(Coolant)
        co.type = "He"
        co.inputs = {"CROSSECTION": topo.ch_area[j], "HYDIAMETER": topo.ch_dh[j]}
        # key order of a running simulation: p, T from the initialisation, density, velocity, then the aliases
        co.dict_node_pt = {"pressure": x[:, M + j].copy(), "temperature": x[:, 2 * M + j].copy()}
        co.dict_node_pt["total_density"] = coolant_mod.PropsSI("Dmass", "T", co.dict_node_pt["temperature"], "P", co.dict_node_pt["pressure"], "He")
        co.dict_node_pt["velocity"] = x[:, j].copy()
        co.dict_node_pt = co._eval_properties(co.dict_node_pt, aliases)
        co.dict_node_pt = co._compute_density_and_mass_flow_rates(co.dict_node_pt)
        ch = object.__new__(Channel)
        ch.inputs = {"HYDIAMETER": topo.ch_dh[j], "FRICTION_MULTIPLIER": model.ch_friction_mult[j], "Roughness": 0.0}
        kv = [("laminar", None), ("turbulent", None), ("total", 0.0)]
        ch.dict_friction_factor = {None: dict(kv), True: dict(kv), False: dict(kv)}
        if model.ch_ifriction[j] == -99:
            ch.turbulent_friction_factor_correlation = ch.user_defined_turbulent_friction_factor
            ch.laminar_friction_factor_correlation = ch.user_defined_laminar_friction_factor
            ch.total_friction_factor_correlation = ch.user_defined_friction_factor
        else:
            ch.turbulent_friction_factor_correlation = ch.rectangular_duct_enea_hts_cicc_trubulent_flow_hole
            ch.laminar_friction_factor_correlation = ch.rectangular_duct_enea_hts_cicc_trubulent_flow_hole
            ch.total_friction_factor_correlation = ch._eval_total_friction_factor_max
        ch.eval_friction_factor(co.dict_node_pt["Reynolds"], nodal=True)
        fluids.append(types.SimpleNamespace(identifier=f"CHAN_{j + 1}", coolant=co, channel=ch))
    bf = np.linspace(sweep.b_field[0, 0, 0], sweep.b_field[0, 0, 1], n)
    tcs = np.linspace(6.0, 7.0, n)
    strand = types.SimpleNamespace(identifier="STR_MIX_1", name="STR_MIX",
                                   dict_node_pt={"temperature": x[:, 3 * M].copy(), "B_field": bf, "T_cur_sharing": tcs})
    jacket = types.SimpleNamespace(identifier="Z_JACKET_1", name="Z_JACKET", dict_node_pt={"temperature": x[:, 3 * M + 1].copy()})
    coll = lambda items, name: types.SimpleNamespace(collection=items, number=len(items), name=name)
    cond.inventory = {"FluidComponent": coll(fluids, "CHAN"), "StrandComponent": coll([strand], "STRAND"),
                      "JacketComponent": coll([jacket], "Z_JACKET"), "StrandStabilizerComponent": coll([], "STR_STAB")}
    out_dir = os.path.join(HERE, "save_properties")
    os.makedirs(out_dir, exist_ok=True)
    save_properties(cond, out_dir)
    np.savez(os.path.join(out_dir, "inputs.npz"), sysvar=x.reshape(-1), zcoord=zcoord, b_field=sweep.b_field[0], tcs=tcs)
    print(sorted(os.listdir(out_dir)))


if __name__ == "__main__":
    main()
