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

    # ---- save_simulation_space (utility_functions/output.py:98-347): two stores of the same conductor, the
    # second with TCS_EVALUATION on, so that both strand layouts and the append branch of Time_sd_actual.tsv exist
    from utility_functions.output import save_simulation_space

    E = n - 1
    jc = np.linspace(2.0e9, 2.5e9, n)
    zg = (zcoord[:-1] + zcoord[1:]) / 2.0
    el = rng.uniform(0.0, 1.0, (2, E, 3))
    gauss = lambda q: {"current_along": el[q, :, 0], "delta_voltage_along": el[q, :, 1],
                       "linear_power_el_resistance": np.stack([el[q, :, 2], np.zeros(E)], axis=1)}
    strand.KIND, strand.operations = "StrandMixedComponent", {"TCS_EVALUATION": False}
    strand.dict_node_pt["J_critical"] = jc
    strand.dict_Gauss_pt, jacket.dict_Gauss_pt = gauss(0), gauss(1)
    htc_nodal = {"ch_ch": {"Open": {"CHAN_1_CHAN_2": rng.uniform(100.0, 200.0, n)}, "Close": {"CHAN_1_CHAN_2": rng.uniform(10.0, 20.0, n)}},
                 "ch_sol": {"CHAN_1_STR_MIX_1": rng.uniform(500.0, 900.0, n), "CHAN_2_Z_JACKET_1": rng.uniform(500.0, 900.0, n)},
                 "sol_sol": {"cond": {"STR_MIX_1_Z_JACKET_1": np.full(n, 500.0)}, "rad": {}}}
    cond.grid_features["zcoord_gauss"] = zg
    cond.grid_input = {"NELEMS": E}
    cond.inventory["SolidComponent"] = coll([strand, jacket], "SOLID")
    cond.Space_save = np.array([0.0, 0.25])
    cond.num_step_save = np.zeros(2, dtype=int)
    cond.i_save = 0
    cond.cond_num_step = 0
    cond.cond_time = [0.0]
    cond.heat_rad_jk, cond.heat_exchange_jk_env = {}, {}
    cond.dict_node_pt = {"HTC": htc_nodal}
    out2 = os.path.join(HERE, "save_simulation_space")
    os.makedirs(out2, exist_ok=True)
    for f in os.listdir(out2):
        os.remove(os.path.join(out2, f))
    save_simulation_space(cond, out2, 4)
    cond.cond_num_step, strand.operations["TCS_EVALUATION"] = 25, True
    cond.cond_time.append(0.25)
    cond.heat_rad_jk = {"Z_JACKET_1_Z_JACKET_2": rng.uniform(0.0, 3.0, n)}
    cond.heat_exchange_jk_env = {"Environment_Z_JACKET_1": rng.uniform(0.0, 3.0, n)}
    save_simulation_space(cond, out2, 4)
    np.savez(os.path.join(out2, "inputs.npz"), sysvar=x.reshape(-1), zcoord=zcoord, zcoord_gauss=zg, tcs=tcs, jc=jc, el=el,
             htc_open=htc_nodal["ch_ch"]["Open"]["CHAN_1_CHAN_2"], htc_close=htc_nodal["ch_ch"]["Close"]["CHAN_1_CHAN_2"],
             htc_fs0=htc_nodal["ch_sol"]["CHAN_1_STR_MIX_1"], htc_fs1=htc_nodal["ch_sol"]["CHAN_2_Z_JACKET_1"],
             htc_ss=htc_nodal["sol_sol"]["cond"]["STR_MIX_1_Z_JACKET_1"],
             heat_rad=cond.heat_rad_jk["Z_JACKET_1_Z_JACKET_2"], heat_env=cond.heat_exchange_jk_env["Environment_Z_JACKET_1"])
    print(sorted(os.listdir(out2)), cond.num_step_save, cond.i_save)

    # ---- save_simulation_time (utility_functions/output.py:647-977): 5 calls with a chunk size of 3, TEND reached at
    # the last one -> header creation, a chunk flush, the final flush and Time.tsv
    from utility_functions.output import save_simulation_time

    out3 = os.path.join(HERE, "save_simulation_time")
    os.makedirs(out3, exist_ok=True)
    for f in os.listdir(out3):
        os.remove(os.path.join(out3, f))
    n_calls, tend = 5, 0.4
    states = np.stack([x * (1.0 + 1e-3 * k) for k in range(n_calls)])                  # (calls, N, d)
    extra = rng.uniform(1.0, 2.0, (n_calls, 2, 3, n))                                   # B_field, T_cur_sharing, J_critical per solid
    elg = rng.uniform(0.0, 1.0, (n_calls, 2, E, 4))                                     # the four Gauss-point arrays per solid
    cond.identifier = "CONDUCTOR_1"
    cond.Time_save = np.array([0.0, 3.4, 10.0])
    cond.n_digit_z = 3
    cond.CHUNCK_SIZE = 3
    cond.cond_time = []
    sim = types.SimpleNamespace(num_step=0, transient_input={"TEND": tend},
                                dict_path={"Output_Time_evolution_CONDUCTOR_1_dir": out3})
    for fc in fluids:
        fc.coolant.time_evol = dict(velocity=dict(), pressure=dict(), temperature=dict(), total_density=dict())
        fc.coolant.time_evol_io = {k: list() for k in (
            "time (s)", "velocity_inl (m/s)", "pressure_inl (Pa)", "temperature_inl (K)", "total_density_inl (kg/m^3)",
            "mass_flow_rate_inl (kg/s)", "velocity_out (m/s)", "pressure_out (Pa)", "temperature_out (K)",
            "total_density_out (kg/m^3)", "mass_flow_rate_out (kg/s)")}
        fc.channel.time_evol = dict(friction_factor=dict())
        fc.channel.flow_dir = ("forward", 1.0)
    fluids[1].channel.flow_dir = ("backward", -1.0)
    strand.time_evol = dict(temperature=dict(), B_field=dict(), T_cur_sharing=dict(), J_critical=dict())
    strand.time_evol_gauss = dict(current_along=dict(), delta_voltage_along=dict(), linear_power_el_resistance=dict(),
                                  delta_voltage_along_sum=dict())
    jacket.time_evol = dict(temperature=dict())
    jacket.time_evol_gauss = dict(current_along=dict(), delta_voltage_along=dict(), linear_power_el_resistance=dict())
    for k in range(n_calls):
        xk = states[k]
        for j, fc in enumerate(fluids):
            co = fc.coolant
            co.dict_node_pt = {"pressure": xk[:, M + j].copy(), "temperature": xk[:, 2 * M + j].copy()}
            co.dict_node_pt["total_density"] = coolant_mod.PropsSI("Dmass", "T", co.dict_node_pt["temperature"], "P", co.dict_node_pt["pressure"], "He")
            co.dict_node_pt["velocity"] = xk[:, j].copy()
            co.dict_node_pt = co._eval_properties(co.dict_node_pt, aliases)
            co.dict_node_pt = co._compute_density_and_mass_flow_rates(co.dict_node_pt)
            fc.channel.eval_friction_factor(co.dict_node_pt["Reynolds"], nodal=True)
        for q, sc in enumerate((strand, jacket)):
            sc.dict_node_pt = {"temperature": xk[:, 3 * M + q].copy(), "B_field": extra[k, q, 0], "T_cur_sharing": extra[k, q, 1],
                               "J_critical": extra[k, q, 2]}
            sc.dict_Gauss_pt = {"current_along": elg[k, q, :, 0], "delta_voltage_along": elg[k, q, :, 1],
                                "linear_power_el_resistance": np.stack([elg[k, q, :, 2], np.zeros(E)], axis=1),
                                "delta_voltage_along_sum": elg[k, q, :, 3]}
        cond.cond_time.append(tend * k / (n_calls - 1))
        sim.num_step = k
        save_simulation_time(sim, cond)
    np.savez(os.path.join(out3, "inputs.npz"), states=states.reshape(n_calls, -1), zcoord=zcoord, zcoord_gauss=zg, extra=extra,
             elg=elg, time_save=cond.Time_save, times=np.array(cond.cond_time), tend=tend)
    print(sorted(os.listdir(out3)))


if __name__ == "__main__":
    main()
