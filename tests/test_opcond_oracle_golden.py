"""oracle/props_oracle.c::osc2o_operating_conditions against what the UNMODIFIED reference's own
orchestration produces (tests/golden/make_golden_opcond.py -> opcond_case1_*.npz):
Conductor.eval_transp_coeff, Conductor.__eval_gauss_point_em, SolidComponent.get_heat,
Conductor.__build_heat_source_gauss_pt, run on reference component objects, with CoolProp's
PropsSI replaced by the table lookup that defines the coolant properties of this build.

Tolerance: Gauss means, table values, Gruneisen, heat source and constant HTCs bit-exact;
everything downstream of pow/exp/log10 fits RTOL = 2e-12 (numpy SIMD loops vs glibc libm)."""
import os

import numpy as np
import pytest

from helpers import GOLDEN_DIR, assert_bits
from opensc2_b200.model import Sweep, case1_model, case2_mini_model, case3_mini_model, synthetic_helium_table
from opensc2_b200.problem import case1_topology, case2_mini_topology, case3_mini_topology
from oracle import oracle as O

RTOL = 2e-12


@pytest.mark.parametrize("tag", ["const_t12", "corr_t12", "corr_t5", "mini2_t12", "mini3_t12"])
def test_operating_conditions_oracle_matches_reference(tag):
    case = {"mini2": "case2", "mini3": "case3"}.get(tag[:5], "case1")
    z = np.load(os.path.join(GOLDEN_DIR, f"opcond_{case}_{tag}.npz"))
    topo = {"case1": lambda: case1_topology(1), "case2": case2_mini_topology, "case3": case3_mini_topology}[case]()
    model = {"case1": lambda: case1_model(bool(z["correlations"])), "case2": case2_mini_model, "case3": case3_mini_model}[case]()
    tab = synthetic_helium_table()
    sweep = Sweep(b_field=z["b_field"], eps=z["eps"], q0=z["q0"], q_window=z["q_window"],
                  tcs=z["tcs"] if "tcs" in z.files else None)
    pn = z["peri_ss_nodal"] if "peri_ss_nodal" in z.files else None
    out = O.operating_conditions(topo, model, [tab], sweep, z["sysvar"], int(z["n_nodes"]), float(z["time"]),
                                 peri_ss_nodal=pn)
    for q in (0, 1, 2, 3, 4, 6, 7):                  # v, p, T, rho, c, h, cv: means and table values
        assert_bits(out["fl_gauss"][:, q], z["out_fl_gauss"][:, q], f"fl_gauss[{q}]")
    assert_bits(out["fl_gauss"][:, 5], z["out_fl_gauss"][:, 5], "Gruneisen")
    assert_bits(out["fl_node"], z["out_fl_node"], "fl_node")
    if case == "case3":      # radiative exchange: sigma * (Ti^2 + Tj^2) * (Ti + Tj) / den -- squares only, still exact
        assert np.abs(z["out_sol_q"]).max() > 0
    assert_bits(out["sol_q"], z["out_sol_q"], "Q1/Q2")
    assert_bits(out["sol_gauss"][:, 0], z["out_sol_gauss"][:, 0], "homogenised density")
    assert_bits(out["htc_es"], z["out_htc_es"], "htc_es")
    for key in ("fl_gauss", "sol_gauss", "htc_ff", "htc_fs", "htc_ss"):
        a, b = out[key], z["out_" + key]
        if a.size == 0:
            continue
        err = np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
        assert err <= RTOL, f"{key}: {err:.3e}"
    if tag == "const_t12":
        for key in ("htc_ff", "htc_fs", "htc_ss"):
            assert_bits(out[key], z["out_" + key], key)
    # the fixtures exercise what they claim to
    if case != "case3":
        assert (z["out_sol_q"].max() > 0) == (float(z["time"]) == 12.0)
    assert np.ptp(z["out_sol_gauss"][:, 1, -1]) > 0
