"""The props oracle (oracle/props_oracle.c) against fixtures produced by the UNMODIFIED
reference fits (tests/golden/make_golden_props.py -> props_fits.npz).

Tolerance: the fits are chains of pow/exp/log10; numpy's SIMD loops and glibc's libm differ
in the last ulp, and the 10**(polynomial) forms amplify that by |exponent| * ln(10).
RTOL = 2e-12 is the stated FP64 relative tolerance of every transcendental fit."""
import os

import numpy as np
import pytest

from helpers import GOLDEN_DIR
from oracle import oracle as O

RTOL = 2e-12


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN_DIR, "props_fits.npz"))


def close(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    err = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    err[(a == b)] = 0.0
    assert np.all(err <= RTOL), f"{what}: max rel err {err.max():.3e} at {err.argmax()}"


@pytest.mark.parametrize("rrr", [197.71, 100.0])
def test_copper_conductivity_and_resistivity(g, rrr):
    close(O.fit("k_cu", g["t"], g["b"], p1=rrr), g[f"k_cu_{rrr}"], "k_cu")
    close(O.fit("rhoe_cu", g["t"], g["b"], p1=rrr), g[f"rhoe_cu_{rrr}"], "rhoe_cu")


def test_copper_specific_heat_including_out_of_range_zero(g):
    out = O.fit("cp_cu", g["t"])
    close(out, g["cp_cu"], "cp_cu")
    assert np.count_nonzero(g["cp_cu"] == 0.0) > 0 and np.array_equal(out == 0.0, g["cp_cu"] == 0.0)


def test_nb3sn(g):
    close(O.fit("k_nb3sn", g["tw"]), g["k_nb3sn"], "k_nb3sn")
    tc = O.fit("tc_nb3sn", g["bb"], g["eps"], p1=16.22, p2=32.35)
    close(tc, g["tc_nb3sn"], "tc_nb3sn")
    assert np.count_nonzero(g["tc_nb3sn"] == 0.0) > 0          # B above Bc2: Tc = 0 branch
    close(O.fit("cp_nb3sn", g["tt"], g["tcs"], g["tc_nb3sn"], p1=16.22), g["cp_nb3sn"], "cp_nb3sn")


def test_stainless_steel_whole_array_branches(g):
    close(O.fit("k_ss", g["tw"]), g["k_ss"], "k_ss")
    for name in ("cold", "mid", "warm"):
        close(O.fit("cp_ss", g[f"t_ss_{name}"]), g[f"cp_ss_{name}"], f"cp_ss_{name}")


def test_glass_epoxy(g):
    close(O.fit("k_ge", g["tw"]), g["k_ge"], "k_ge")
    close(O.fit("cp_ge", g["tw"]), g["cp_ge"], "cp_ge")


def test_channel_correlations(g):
    close(O.fit("friction_124", g["re"]) * 0.02, g["f124_total"], "friction 124")
    dh = float(g["hydiameter"])
    close(O.fit("nusselt_1", g["re"], g["pr"]) * g["kf"] / dh, g["htc_1"], "htc corr 1")
    close(O.fit("nusselt_212", g["re"]) * g["kf"] / dh, g["htc_212"], "htc corr 212")


def test_aluminium_whole_array_range_check_and_re123(g):
    close(O.fit("k_al", g["t_al"]), g["k_al"], "k_al")
    close(O.fit("cp_al", g["t_al"]), g["cp_al"], "cp_al")
    assert np.all(g["k_al_bad"] == 0.0) and np.all(g["cp_al_bad"] == 0.0)        # what the reference does
    assert np.all(O.fit("k_al", g["t_al_bad"]) == 0.0) and np.all(O.fit("cp_al", g["t_al_bad"]) == 0.0)
    close(O.fit("k_re123", g["t_re"]), g["k_re123"], "k_re123")
    close(O.fit("cp_re123", g["t_re"]), g["cp_re123"], "cp_re123")


def test_colebrook_array_wide_iteration(g):
    out = O.fit("friction_122", g["re"], p1=float(g["roughness"]), p2=float(g["hydiameter"])) * 0.02
    close(out, g["f122_total"], "friction 122")


def test_stack_layer_materials_against_the_reference_fits():
    """Silver, Hastelloy C276, Sn60Pb40 (StackComponent layers): oracle fits vs the reference functions,
    including the np.piecewise edge values (T = 13, 35, 60, 285, 300, 1234 K ...)."""
    import os
    from helpers import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, "props_stack.npz"))
    for key in ("ag", "hc276", "sn60pb40"):
        for q in ("k", "cp"):
            got = O.fit(f"{q}_{key}", g["t"])
            ref = g[f"{q}_{key}"]
            err = np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300))
            assert err <= 2e-12, f"{q}_{key}: {err:.3e}"
            assert np.array_equal(got == 0.0, ref == 0.0)


def test_stack_component_homogenisation_against_the_reference():
    """StackComponent.stack_density / stack_isobaric_specific_heat / stack_thermal_conductivity run by the
    reference on a five-layer tape vs the oracle's operating conditions for a SOLID_STACK solid."""
    from helpers import stack_case
    topo, model, tab, arr, sweep, g = stack_case()
    ref = O.operating_conditions(topo, model, [tab], sweep, arr.sysvar, arr.n_nodes, 5.0)
    sg = ref["sol_gauss"]            # (B, 3, S, E)
    for q, key in enumerate(("stack_rho", "stack_cp", "stack_k")):
        for e in range(arr.n_nodes - 1):
            err = np.max(np.abs(sg[:, q, 0, e] - g[key]) / np.abs(g[key]))
            assert err <= 2e-12, f"{key}: {err:.3e}"


@pytest.mark.parametrize("ifr", [100, 101, 110, 111, 112, 113, 114, 115, 118, 119, 120, 121, 200, 201, 202, 203, 204, 205, 206, 207, 209, 210,
                                 211])
def test_more_friction_laws_against_the_reference_channel(ifr):
    """IFRICTION 110 / 111 / 112 / 121 and the bundle laws 204 / 205 (Katheder), 206 / 209 (Darcy-Forchheimer)
    through the reference's Channel.eval_friction_factor (laminar, turbulent
    and total slots, multiplier 0.02) vs the oracle, including negative Reynolds numbers, the mask gaps of the
    LHC recipe at Re = 2e3 / 4e3 and the laminar cut at Re < 1e-5."""
    import os
    from helpers import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, "props_stack.npz"))
    re = g["re"]
    from opensc2_b200.model import duct_laminar_coefficient
    y = float(g["crossection"]) if ifr == 200 else float(g["roughness"])        # the fit's second argument: CROSSECTION for 200,
    if ifr == 119:                                                                # the laminar numerator for 119
        y = duct_laminar_coefficient(float(g["side_ratio"]))
    got = O.fit("friction_simple", re, np.full_like(re, y), np.full_like(re, float(g["hydiameter"])), p1=ifr,
                p2=float(g["void_fraction"])) * (1.0 if ifr == 211 else 0.02)        # 211 forces FRICTION_MULTIPLIER to 1
    ref = g[f"f{ifr}_total"]
    # the host-side numpy restatement (gen_flow, writers) against the same reference vectors
    from opensc2_b200.correlations import friction_factor
    host = friction_factor(ifr, 0.02, re, float(g["hydiameter"]), float(g["roughness"]), float(g["void_fraction"]), float(g["crossection"]),
                           side_ratio=float(g["side_ratio"]))
    assert np.max(np.abs(host - ref) / np.maximum(np.abs(ref), 1e-300)) <= 2e-12, f"host IFRICTION {ifr}"
    err = np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300))
    assert err <= 2e-12, f"IFRICTION {ifr}: {err:.3e}"


def test_niobium_titanium_against_the_reference_fits():
    """NbTi: thermal conductivity, Tc(B) and cp(T, B, Tcs, Tc) -- the four-term fit above 1 K and the
    field-indexed superconducting-state table mixed between Tcs and Tc below it."""
    import os
    from helpers import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, "props_stack.npz"))
    bc20m, tc0m = g["nbti_bc20m_tc0m"]
    checks = {
        "k_nbti": (O.fit("k_nbti", g["nbti_t"]), g["k_nbti"]),
        "tc_nbti": (O.fit("tc_nbti", g["nbti_b"], p1=bc20m, p2=tc0m), g["nbti_tc"]),
        "cp_nbti": (O.fit("cp_nbti", g["nbti_t"], g["nbti_b"], g["nbti_tcs"], out_init=g["nbti_tc"]), g["cp_nbti"]),
    }
    for key, (got, ref) in checks.items():
        err = np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300))
        assert err <= 2e-12, f"{key}: {err:.3e}"
    assert np.any(g["nbti_tc"] == 0.0) and np.any(g["nbti_t"] <= 1.0)


@pytest.mark.parametrize("corr", [119, 120, 121, 211])
def test_more_steady_state_htc_correlations_against_the_reference_channel(corr):
    """Flag_htc_steady_corr 119 / 120 / 121 (Dittus-Boelter with the lower limits of channel.py:1171-1243) and 211,
    through the reference's Channel.eval_steady_state_htc: h = Nu k / Dh."""
    import os
    from helpers import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, "props_stack.npz"))
    dh = float(g["hydiameter"])
    got = O.fit("nusselt_more", np.abs(g["re"]), g["pr"], p1=corr, p2=dh) * g["kf"] / dh
    ref = g[f"htc_{corr}"]
    err = np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300))
    assert err <= 2e-12, f"Flag_htc_steady_corr {corr}: {err:.3e}"


def test_magnesium_diboride_and_all_densities_against_the_reference():
    """MgB2 conductivity (T, B; the copper NIST form at rrr = 100, as strand_mixed_component.py calls it) and
    cp(T) with its out-of-range zeros; the constant density of every material id."""
    import os
    from helpers import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, "props_stack.npz"))
    for key, got in (("k_mgb2", O.fit("k_mgb2", g["mgb2_t"], g["mgb2_b"])), ("cp_mgb2", O.fit("cp_mgb2", g["mgb2_t"]))):
        ref = g[key]
        err = np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300))
        assert err <= 2e-12, f"{key}: {err:.3e}"
        assert np.array_equal(got == 0.0, ref == 0.0)
    for mat, rho in enumerate(g["densities"]):
        assert O.fit("density", np.ones(1), p1=mat)[0] == rho, mat


@pytest.mark.parametrize("ifr", list(range(102, 110)))
def test_newton_hole_friction_laws_against_the_reference_channel(ifr):
    """IFRICTION 102..109 (turbulent_friction_with_newton_hole + _eval_ft_newton, channel.py:190-222, 349-390): the
    fixed-point iteration runs on the WHOLE Reynolds array until the largest relative change is below 1e-2, so every
    element gets the array-wide number of iterations; 106 takes the interpolated total, the others the maximum.  Oracle
    (array loop) and host numpy restatement against the reference's Channel.eval_friction_factor."""
    import os
    from helpers import GOLDEN_DIR
    from opensc2_b200.correlations import friction_factor
    g = np.load(os.path.join(GOLDEN_DIR, "props_stack.npz"))
    re, dh, ref = g["re"], float(g["hydiameter"]), g[f"f{ifr}_total"]
    # the array loop itself, on a vector with Reynolds numbers below ~100 where the iteration wanders chaotically: the
    # oracle (same libm, same loop) lands on the reference's iteration count and values
    got = O.fit("friction_hole_newton", re, p1=ifr, p2=dh) * 0.02
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)) <= 2e-12, f"oracle IFRICTION {ifr}"
    # where every element converges (Re >= 1e3) the result does not hang on the last bit of pow(): oracle and the host
    # numpy restatement against the reference
    re, ref = g["re_regular"], g[f"f{ifr}_total_regular"]
    got = O.fit("friction_hole_newton", re, p1=ifr, p2=dh) * 0.02
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)) <= 2e-12, f"oracle IFRICTION {ifr}, regular"
    host = friction_factor(ifr, 0.02, re, dh)
    assert np.max(np.abs(host - ref) / np.maximum(np.abs(ref), 1e-300)) <= 2e-12, f"host IFRICTION {ifr}"

