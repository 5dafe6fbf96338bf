"""Generate tests/golden/props_fits.npz by running the UNMODIFIED reference modules
(properties_of_materials/*.py, channel.py, strand_mixed_component.py, jacket_component.py).

Build-container only (needs /root/reference):  python tests/golden/make_golden_props.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import ref_harness as H  # noqa: E402


def main():
    assert H.reference_available()
    H._import_reference()
    from properties_of_materials.copper import (electrical_resistivity_cu_nist, isobaric_specific_heat_cu_nist,
                                                thermal_conductivity_cu_nist)
    from properties_of_materials.glass_epoxy import isobaric_specific_heat_ge, thermal_conductivity_ge
    from properties_of_materials.niobium3_tin import (critical_temperature_nb3sn, isobaric_specific_heat_nb3sn,
                                                      thermal_conductivity_nb3sn)
    from properties_of_materials.stainless_steel import isobaric_specific_heat_ss, thermal_conductivity_ss
    from channel import Channel

    rng = np.random.default_rng(20261019)
    n = 400
    out = {}
    t = np.concatenate([rng.uniform(3.0, 30.0, n // 2), rng.uniform(30.0, 320.0, n // 2 - 4), [4.0, 300.0, 3.99, 300.5]])
    b = np.concatenate([rng.uniform(0.0, 13.0, n - 40), np.zeros(40)])
    rng.shuffle(b)
    out["t"], out["b"] = t, b
    for rrr in (197.71, 100.0):
        out[f"k_cu_{rrr}"] = thermal_conductivity_cu_nist(t, b, rrr)
        out[f"rhoe_cu_{rrr}"] = electrical_resistivity_cu_nist(t, b, rrr)
    out["cp_cu"] = isobaric_specific_heat_cu_nist(t)
    tw = np.concatenate([rng.uniform(1.0, 40.0, n // 2), rng.uniform(40.0, 800.0, n // 2)])
    out["tw"] = tw
    out["k_nb3sn"] = thermal_conductivity_nb3sn(tw)
    out["k_ss"] = thermal_conductivity_ss(tw)
    out["k_ge"] = thermal_conductivity_ge(tw)
    out["cp_ge"] = isobaric_specific_heat_ge(tw)
    # SS specific heat: the branch depends on max(T) of the whole array
    for name, arr in (("cold", rng.uniform(0.2, 1.0, 50)), ("mid", rng.uniform(1.0, 12.5, 50)), ("warm", tw)):
        out[f"t_ss_{name}"] = arr
        out[f"cp_ss_{name}"] = isobaric_specific_heat_ss(arr)
    # Nb3Sn: Tc(B, eps), cp(T, Tcs, Tc)
    bb = np.concatenate([rng.uniform(0.0, 13.0, n - 20), rng.uniform(28.0, 40.0, 20)])
    eps = rng.uniform(-0.008, 0.002, n)
    out["bb"], out["eps"] = bb, eps
    tc0m, bc20m = 16.22, 32.35
    tc = critical_temperature_nb3sn(bb, eps, tc0m, bc20m)
    out["tc_nb3sn"] = tc
    tcs = np.where(rng.uniform(size=n) < 0.8, tc * rng.uniform(0.3, 1.0, n), tc * 1.05)
    tt = np.concatenate([rng.uniform(3.0, 25.0, n - 50), rng.uniform(25.0, 1100.0, 50)])
    out["tt"], out["tcs"] = tt, tcs
    out["cp_nb3sn"] = isobaric_specific_heat_nb3sn(tt, tcs, tc, tc0m)
    # channel correlations through the reference Channel methods on a bare object
    ch = types.SimpleNamespace()
    for name in ("blasius_formula", "rectangular_duct_enea_hts_cicc_trubulent_flow_hole", "_quantities_initialization",
                 "_eval_total_friction_factor_max", "eval_friction_factor", "dittus_boelter",
                 "dittus_boelter_nusselt_lower_limit", "rectangular_duct_enea_hts_cicc_htc",
                 "_initialize_nusselt_and_htc", "eval_steady_state_htc", "_eval_nusselt_lower_limit_119"):
        setattr(ch, name, types.MethodType(getattr(Channel, name), ch))
    ch.inputs = {"FRICTION_MULTIPLIER": 0.02, "HYDIAMETER": 2.9115e-4, "Flag_htc_steady_corr": 1,
                 "Side": 1.0e-3, "Base": 1.0e-3, "Height": 1.0e-3}
    ch.dict_friction_factor = {True: {"laminar": None, "turbulent": None, "total": 0.0}}
    ch.laminar_friction_factor_correlation = ch.rectangular_duct_enea_hts_cicc_trubulent_flow_hole
    ch.turbulent_friction_factor_correlation = ch.rectangular_duct_enea_hts_cicc_trubulent_flow_hole
    ch.total_friction_factor_correlation = ch._eval_total_friction_factor_max
    re = np.concatenate([rng.uniform(50.0, 1.0e4, n // 2), rng.uniform(1.0e4, 5.0e5, n // 2)])
    pr = rng.uniform(0.5, 2.0, n)
    kf = rng.uniform(0.01, 0.03, n)
    out["re"], out["pr"], out["kf"] = re, pr, kf
    ch.eval_friction_factor(re.copy(), nodal=True)
    out["f124_total"] = ch.dict_friction_factor[True]["total"]
    ch.dict_nusselt, ch.dict_htc_steady = {True: None}, {True: None}
    prop = {"Reynolds": re, "Prandtl": pr, "total_thermal_conductivity": kf}
    try:
        ch.nusselt_correlation = ch.dittus_boelter_nusselt_lower_limit
        ch.eval_steady_state_htc(prop, nodal=True)
        out["htc_1"] = ch.dict_htc_steady[True].copy()
    except Exception as e:                     # _eval_nusselt_lower_limit_119 needs geometry keys
        raise
    ch.inputs["Flag_htc_steady_corr"] = 212
    ch.nusselt_correlation = ch.rectangular_duct_enea_hts_cicc_htc
    ch.eval_steady_state_htc(prop, nodal=True)      # the call of conductor.py:4842-4844: nn stays 0.3
    out["htc_212"] = ch.dict_htc_steady[True].copy()
    out["hydiameter"] = np.array(ch.inputs["HYDIAMETER"])
    # IFRICTION 122: Colebrook started from Haaland, array-wide convergence (channel.py:700-756); the
    # laminar slot calls the same function (channel.py:109-112), total = max(0, turbulent) * multiplier
    for name in ("haaland_equation", "colebrook_formula"):
        setattr(ch, name, types.MethodType(getattr(Channel, name), ch))
    ch.inputs["Roughness"] = 2.0e-6
    ch.laminar_friction_factor_correlation = ch.colebrook_formula
    ch.turbulent_friction_factor_correlation = ch.colebrook_formula
    ch.total_friction_factor_correlation = ch._eval_total_friction_factor_max
    ch.dict_friction_factor = {True: {"laminar": None, "turbulent": None, "total": 0.0}}
    ch.eval_friction_factor(re.copy(), nodal=True)
    out["f122_total"] = ch.dict_friction_factor[True]["total"]
    out["roughness"] = np.array(ch.inputs["Roughness"])
    # aluminium (whole-array range check) and RE123
    from properties_of_materials.aluminium import isobaric_specific_heat_al, thermal_conductivity_al
    from properties_of_materials.rare_earth_123 import isobaric_specific_heat_re123, thermal_conductivity_re123
    t_al = np.concatenate([rng.uniform(2.0, 12.0, 150), rng.uniform(12.0, 60.0, 150), rng.uniform(60.0, 1000.0, 100),
                           [2.0, 4.0, 6.0, 11.0578562, 53.2685055, 300.0, 1000.0]])
    out["t_al"] = t_al
    out["k_al"], out["cp_al"] = thermal_conductivity_al(t_al), isobaric_specific_heat_al(t_al)
    t_al_bad = t_al.copy()
    t_al_bad[5] = 1.5                      # one element below 2 K: the reference returns zeros for the WHOLE array
    out["t_al_bad"] = t_al_bad
    out["k_al_bad"], out["cp_al_bad"] = thermal_conductivity_al(t_al_bad), isobaric_specific_heat_al(t_al_bad)
    t_re = np.concatenate([rng.uniform(2.0, 70.0, 200), rng.uniform(70.0, 350.0, 200)])
    out["t_re"] = t_re
    out["k_re123"], out["cp_re123"] = thermal_conductivity_re123(t_re), isobaric_specific_heat_re123(t_re)
    path = os.path.join(HERE, "props_fits.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
