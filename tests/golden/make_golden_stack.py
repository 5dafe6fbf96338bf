"""Generate tests/golden/props_stack.npz by running the UNMODIFIED reference: the silver, Hastelloy C276 and
Sn60Pb40 fits (properties_of_materials/) and the StackComponent homogenisers (stack_component.py:398-476) on
a bare StackComponent object (its constructor needs openpyxl); and the friction laws IFRICTION 100-115, 118,
120, 121, 200-207, 209-211 through Channel.eval_friction_factor on a bare object wired like Channel.__init__ (channel.py:43-169).

Build-container only (needs /root/reference):  python tests/golden/make_golden_stack.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import ref_harness as H  # noqa: E402

# tape of the reference's StackComponent: (material key, layer thickness in m), in tape order
TAPE = (("ag", 2.0e-6), ("cu", 40.0e-6), ("hc276", 50.0e-6), ("re123", 1.0e-6), ("sn60pb40", 10.0e-6))
RRR = 100.0


def main():
    assert H.reference_available()
    H._import_reference()
    import stack_component as SC

    rng = np.random.default_rng(20261020)
    n = 600
    t = np.concatenate([rng.uniform(0.5, 70.0, n // 3), rng.uniform(40.0, 360.0, n // 3), rng.uniform(250.0, 1300.0, n // 3 - 12),
                        [1.0, 13.0, 35.0, 50.0, 60.0, 100.0, 273.0, 285.0, 300.0, 350.0, 1234.0, 4.0]])
    out = {"t": t}
    for key in ("ag", "hc276", "sn60pb40"):
        out[f"k_{key}"] = SC.THERMAL_CONDUCTIVITY_FUNC[key](t)
        out[f"cp_{key}"] = SC.ISOBARIC_SPECIFIC_HEAT_FUNC[key](t)
        out[f"rho_{key}"] = SC.DENSITY_FUNC[key](t)
    # homogenised stack through the reference's own methods
    st = object.__new__(SC.StackComponent)
    keys = [k for k, _ in TAPE]
    st.inputs = {"Material_number": len(TAPE), "RRR": RRR}
    st.material_thickness = np.array([s for _, s in TAPE])
    st.tape_thickness = st.material_thickness.sum()
    st.density_function = np.array([SC.DENSITY_FUNC[k] for k in keys])
    st.isobaric_specific_heat_function = np.array([SC.ISOBARIC_SPECIFIC_HEAT_FUNC[k] for k in keys])
    st.thermal_conductivity_function = np.array([SC.THERMAL_CONDUCTIVITY_FUNC[k] for k in keys])
    ts = np.concatenate([rng.uniform(4.0, 100.0, 200), rng.uniform(20.0, 320.0, 100)])
    bs = rng.uniform(0.0, 12.0, ts.size)
    prop = {"temperature": ts, "B_field": bs}
    out["stack_t"], out["stack_b"] = ts, bs
    out["stack_rho"] = st.stack_density(prop)
    out["stack_cp"] = st.stack_isobaric_specific_heat(prop)
    out["stack_k"] = st.stack_thermal_conductivity(prop)
    out["stack_thickness"] = st.material_thickness
    # friction laws through the reference's own three-slot evaluation (laminar, turbulent, total)
    import types
    from channel import Channel
    re = np.concatenate([rng.uniform(1.0, 2.0e3, 100), rng.uniform(2.0e3, 4.0e3, 60), rng.uniform(4.0e3, 3.0e4, 100),
                         rng.uniform(3.0e4, 1.0e6, 100), [2.0e3, 4.0e3, 3.0e4, 1.0e-6, 0.5]])
    re = re * np.where(rng.uniform(size=re.size) < 0.2, -1.0, 1.0)          # backward flow: Reynolds made positive inside
    out["re"] = re
    out["roughness"], out["hydiameter"] = np.array(2.0e-6), np.array(2.499e-3)
    wiring = {110: ("general_laminar_correlation", "blasius_formula"),
              111: ("general_laminar_correlation", "bhatti_shah_correlation_turbulent_flow_hole"),
              112: ("general_laminar_correlation", "lhc_magnets_recipe_hole"),
              121: ("haaland_equation", "haaland_equation"),
              204: ("general_laminar_correlation", "katheder_correlation_bundle"),
              205: ("general_laminar_correlation", "katheder_correlation_bundle"),
              206: ("general_laminar_correlation", "darcy_forcheimer_porus_medium_correlation_bundle"),
              209: ("general_laminar_correlation", "darcy_forcheimer_porus_medium_correlation_bundle"),
              113: ("general_laminar_correlation", "memorandum_bessette_iter_cs_spiral_hole_y2015"),
              114: ("general_laminar_correlation", "csi_hole_ls_best_fit_icec_hole_y2016"),
              115: ("general_laminar_correlation", "tronza_iter_tf_hole_correlation"),
              118: ("general_laminar_correlation", "flat_spiral_from_cfd_hole"),
              200: ("general_laminar_correlation", "tronza_iter_tf_correlation_bundle"),
              201: ("general_laminar_correlation", "nicollet_general_correlation_bundle"),
              202: ("general_laminar_correlation", "wanner_correlation_bundle"),
              203: ("general_laminar_correlation", "kstar_correlation_bundle"),
              207: ("general_laminar_correlation", "lhc_poncet_correlation_bundle"),
              210: ("general_laminar_correlation", "demo_tf_hts_correlation_bundle"),
              211: ("hts_cl_correlation_bundle", "hts_cl_correlation_bundle"),
              100: ("general_laminar_correlation", "iter_conductor_hole"),
              101: ("general_laminar_correlation", "iter_conductor_hole"),
              119: ("rectangular_duct_demo_common_memo_laminar_flow_hole", "duct_demo_common_memo_turbulent_flow_hole"),
              120: ("triangular_duct_demo_common_memo_laminar_flow_hole", "duct_demo_common_memo_turbulent_flow_hole")}
    out["side_ratio"] = np.array(2.0e-3 / 5.0e-3)
    wiring.update({k: ("general_laminar_correlation", "turbulent_friction_with_newton_hole") for k in range(102, 110)})
    out["crossection"] = np.array(3.7e-4)
    out["void_fraction"] = np.array(0.33)
    rng2 = np.random.default_rng(20261020)           # its own generator: the vectors drawn further down stay what they were
    out["re_regular"] = np.concatenate([rng2.uniform(1.0e3, 4.0e3, 60), rng2.uniform(4.0e3, 1.0e6, 140)]) * np.where(rng2.uniform(size=200) < 0.2, -1.0, 1.0)
    for ifr, (lam, turb) in wiring.items():
        ch = types.SimpleNamespace()
        for name in ("_quantities_initialization", "_eval_total_friction_factor_max", "_eval_total_friction_factor_L_lim_T",
                     "eval_friction_factor", "_eval_jj_and_kk", "_eval_correction_diameter_hole", "_eval_ft_newton", lam, turb):
            setattr(ch, name, types.MethodType(getattr(Channel, name), ch))
        ch.inputs = {"FRICTION_MULTIPLIER": 0.02, "HYDIAMETER": 2.499e-3, "Roughness": 2.0e-6, "IFRICTION": ifr, "VOID_FRACTION": 0.33, "CROSSECTION": 3.7e-4,
                     "SIDE1": 5.0e-3, "SIDE2": 2.0e-3}
        ch.dict_friction_factor = {True: {"laminar": None, "turbulent": None, "total": 0.0}}
        ch.laminar_friction_factor_correlation = getattr(ch, lam)
        ch.turbulent_friction_factor_correlation = getattr(ch, turb)
        ch.total_friction_factor_correlation = (ch._eval_total_friction_factor_L_lim_T if ifr in (106, 119, 120)
                                                else ch._eval_total_friction_factor_max)      # channel.py:116-132
        ch.eval_friction_factor(re.copy(), nodal=True)
        out[f"f{ifr}_total"] = np.array(ch.dict_friction_factor[True]["total"])
        if 102 <= ifr <= 109:
            # the fixed-point iteration of these laws is chaotic below Re ~ 100 (it never settles; the array-wide loop of the
            # reference ends when the wandering element happens to move by less than 1 % in the same sweep as the others):
            # a second vector on Reynolds numbers where every element converges monotonically, i.e. where the result does not
            # depend on the last bit of pow()
            ch.dict_friction_factor = {True: {"laminar": None, "turbulent": None, "total": 0.0}}
            ch.eval_friction_factor(out["re_regular"].copy(), nodal=True)
            out[f"f{ifr}_total_regular"] = np.array(ch.dict_friction_factor[True]["total"])
    # steady-state heat transfer: Flag_htc_steady_corr 119 / 120 / 121 / 211 through Channel.eval_steady_state_htc
    pr = rng.uniform(0.5, 2.0, re.size)
    kf = rng.uniform(0.01, 0.03, re.size)
    rea = np.abs(re)
    out["pr"], out["kf"] = pr, kf
    for corr, method in ((119, "dittus_boelter_nusselt_lower_limit"), (120, "dittus_boelter_nusselt_lower_limit"),
                         (121, "dittus_boelter_nusselt_lower_limit"), (211, "correlation_211")):
        ch = types.SimpleNamespace()
        for name in ("dittus_boelter", "dittus_boelter_nusselt_lower_limit", "correlation_211", "_initialize_nusselt_and_htc",
                     "eval_steady_state_htc", "_eval_nusselt_lower_limit_119"):
            setattr(ch, name, types.MethodType(getattr(Channel, name), ch))
        ch.inputs = {"HYDIAMETER": 2.499e-3, "Flag_htc_steady_corr": corr}
        ch.dict_nusselt, ch.dict_htc_steady = {True: None}, {True: None}
        ch.nusselt_correlation = getattr(ch, method)
        ch.eval_steady_state_htc({"Reynolds": rea, "Prandtl": pr, "total_thermal_conductivity": kf}, nodal=True)   # conductor.py:4842-4844
        out[f"htc_{corr}"] = np.array(ch.dict_htc_steady[True])
    # NbTi (superconductor of a StrandMixedComponent): conductivity, Tc(B), cp(T, B, Tcs, Tc) incl. the T <= 1 K table
    from properties_of_materials.niobium_titanium import (critical_temperature_nbti, isobaric_specific_heat_nbti,
                                                          thermal_conductivity_nbti)
    tn = np.concatenate([rng.uniform(0.2, 1.0, 120), rng.uniform(1.0, 30.0, 200), rng.uniform(30.0, 1100.0, 76), [1.0, 50.0, 750.0, 1000.0]])
    bn = np.concatenate([rng.uniform(0.0, 8.0, 380), rng.uniform(13.0, 16.0, 16), [2.0, 3.0, 7.0, 14.5]])
    bc20m, tc0m = 14.5, 9.03
    tcn = critical_temperature_nbti(bn, bc20m, tc0m)
    tcsn = np.where(rng.uniform(size=tn.size) < 0.8, tcn * rng.uniform(0.05, 1.0, tn.size), tcn * 1.05)
    out["nbti_t"], out["nbti_b"], out["nbti_tc"], out["nbti_tcs"] = tn, bn, tcn, tcsn
    out["nbti_bc20m_tc0m"] = np.array([bc20m, tc0m])
    out["k_nbti"] = thermal_conductivity_nbti(tn)
    out["cp_nbti"] = isobaric_specific_heat_nbti(tn.copy(), bn, tcsn, tcn)
    # MgB2: conductivity (T, B) as strand_mixed_component.py calls it (default rrr), cp(T) incl. out-of-range zeros
    from properties_of_materials.magnesium_diboride import isobaric_specific_heat_mgb2, thermal_conductivity_mgb2
    tm = np.concatenate([rng.uniform(4.0, 40.0, 200), rng.uniform(40.0, 300.0, 96), [3.9, 4.0, 300.0, 310.0]])
    bm = np.concatenate([rng.uniform(0.0, 6.0, 260), np.zeros(40)])
    out["mgb2_t"], out["mgb2_b"] = tm, bm
    out["k_mgb2"] = thermal_conductivity_mgb2(tm, bm)
    out["cp_mgb2"] = isobaric_specific_heat_mgb2(tm)
    # constant densities of every material id of include/opensc2_b200.h, in the order of its enum
    import strand_mixed_component as SM
    import jacket_component as JK
    one = np.ones(1)
    out["densities"] = np.array([SC.DENSITY_FUNC["cu"](one)[0], SM.DENSITY_FUNC["nb3sn"](one)[0], JK.DENSITY_FUNC["ss"](one)[0],
                                 JK.DENSITY_FUNC["ge"](one)[0], SC.DENSITY_FUNC["al"](one)[0], SC.DENSITY_FUNC["re123"](one)[0],
                                 SC.DENSITY_FUNC["ag"](one)[0], SC.DENSITY_FUNC["hc276"](one)[0], SC.DENSITY_FUNC["sn60pb40"](one)[0],
                                 SM.DENSITY_FUNC["nbti"](one)[0], SM.DENSITY_FUNC["mgb2"](one)[0]])
    np.savez_compressed(os.path.join(HERE, "props_stack.npz"), **out)
    print("wrote props_stack.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
