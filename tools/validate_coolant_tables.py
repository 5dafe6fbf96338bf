"""Validation report of the coolant table generator (opensc2_b200/eos) against CoolProp 6.4.1 values -- the ones
the reference itself wrote into its golden `TDD_examples/*/Solution/CHAN_*.tsv` files.  Build container only.

What the golden channel files hold (established here, and the reason the columns can be used as pins):
  * `total_density` = PropsSI("Dmass") at the (p, T) of the SAME row (post-step recompute, coolant.py:247-263);
  * every other property column = PropsSI at the INITIAL state of the run (uniform inlet temperature, linear
    pressure between the inlet pressure and the outlet pressure gen_flow computed): the revision that produced the
    goldens never refreshed the nodal property arrays.  Evidence: with T = 100 K and p linear between the golden's
    own p(0) = 260 115.74 Pa and PREINL = 3.0e5 Pa, all nine nitrogen columns agree with the generator to <= 6e-8.
Usage:  python tools/validate_coolant_tables.py  [> profiles/r2_coolant_table_validation.txt]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opensc2_b200.deck import load_deck  # noqa: E402
from opensc2_b200.eos import build_table, interpolation_error  # noqa: E402
from opensc2_b200.eos import nitrogen as N2  # noqa: E402
from opensc2_b200.initial_state import gen_flow  # noqa: E402

CASE3 = "/root/reference/TDD_examples/CASE_3_HTS_HVDC"
COLS17 = {"rho": 3, "beta": 5, "kappa": 6, "prandtl": 7, "mu": 8, "h": 9, "cp": 10, "cv": 11, "w": 12, "k": 13}


def main():
    a = np.loadtxt(os.path.join(CASE3, "Solution", "CHAN_2.tsv"), skiprows=1)
    z, p, T = a[:, 0], a[:, 1], a[:, 2]
    print("== nitrogen (Span et al. 2000; Lemmon & Jacobsen 2004) vs golden CASE_3 CHAN_2.tsv, 201 rows ==")
    s = N2.state(T, p)
    print(f"Dmass at the row's own (p, T), 100.0-132.8 K, 2.60-3.07 bar: max |rel err| = {np.max(np.abs(s['rho'] / a[:, 3] - 1)):.2e}")
    p0 = p[0] + (3.0e5 - p[0]) * z / z[-1]          # initial pressure line (backward flow: inlet at x = L)
    s0 = N2.state(np.full_like(z, 100.0), p0)
    for k, c in COLS17.items():
        if k == "rho":
            continue
        err = s0[k] / a[:, c] - 1
        print(f"{k:8s} at the initial state (100 K, {p0[0]:.2f}..{p0[-1]:.2f} Pa): max |rel err| = {np.max(np.abs(err)):.2e}")
    print(f"Gruneisen column (beta / kappa / cv / rho): max |rel err| = "
          f"{np.max(np.abs(s0['beta'] / s0['kappa'] / s0['cv'] / s0['rho'] / a[:, 15] - 1)):.2e}")
    print(f"enthalpy datum: golden - generator = {np.mean(a[:, 9] - s0['h']):+.4f} J/kg (std {np.std(a[:, 9] - s0['h']):.1e}) "
          "-> CoolProp keeps the datum of the EOS paper for nitrogen")
    print("\n== bilinear interpolation error of the device table vs the equation of state (4000 random interior points) ==")
    for n_p, n_t in ((65, 129), (201, 281), (401, 561), (801, 1121)):
        tab = build_table(N2.state, (2.0e5, 4.0e5), (90.0, 160.0), n_p, n_t)
        e = interpolation_error(tab, N2.state)
        worst = max(e, key=e.get)
        d = load_deck(CASE3, intial_override=2)
        bc = gen_flow(d.topology, d.model, [tab, tab], d.fl_bc.copy(), d.length, np.zeros(0))
        print(f"nitrogen 2-4 bar x 90-160 K, {n_p} x {n_t} ({tab.props.nbytes / 1e6:.1f} MB): max rel err {e[worst]:.1e} ({worst}), "
              f"rho {e['rho']:.1e}; gen_flow outlet pressure of CHAN_2 from this table: {bc[1, 1]:.5f} Pa "
              f"(golden p(0) = {p[0]:.5f}, rel {bc[1, 1] / p[0] - 1:+.1e})")
    print("\n== helium ==")
    print("No validated helium equation of state could be produced in this image: CoolProp 6.4.1 (Ortiz-Vega helium EOS,\n"
          "Arp/McCarty/Friend viscosity, Hands/Arp conductivity) is absent and un-vendored, there is no network, and a\n"
          "coefficient set reproduced from memory FAILED this same validation (density at 4.5 K / 6 bar: 102.8 vs the\n"
          "golden 139.607 kg/m^3), so it is not shipped.  Helium tables stay an INPUT of the library (generate them with\n"
          "CoolProp where it exists: opensc2_b200.eos.build_table(state, ...) takes any state(T, p) callable); the\n"
          "helium numbers of bench.py and of the transient tests use model.synthetic_helium_table / helium_gas_table and\n"
          "say so.  What the goldens pin for helium without an EOS: see tests/test_case_goldens.py.")


if __name__ == "__main__":
    main()
