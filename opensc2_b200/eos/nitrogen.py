"""Nitrogen as CoolProp 6.4.1 models it (the second coolant of TDD_examples/CASE_3_HTS_HVDC):
  * equation of state: Span, Lemmon, Jacobsen, Wagner, Yokozeki, "A Reference Equation of State for the
    Thermodynamic Properties of Nitrogen for Temperatures from 63.151 to 1000 K and Pressures to 2200 MPa",
    J. Phys. Chem. Ref. Data 29 (2000) 1361 -- 36-term residual part, Planck-Einstein ideal part;
  * viscosity and thermal conductivity: Lemmon, Jacobsen, "Viscosity and Thermal Conductivity Equations for
    Nitrogen, Oxygen, Argon, and Air", Int. J. Thermophys. 25 (2004) 21 (dilute gas from the Lennard-Jones collision
    integral + residual part; conductivity with the simplified Olchowy-Sengers critical enhancement).
The coefficients are the published ones.  Validation: `tools/validate_coolant_tables.py` against every
`total_*` column of the reference's golden CASE_3 CHAN_2.tsv (201 points, 100-133 K, 2.6-3.1 bar).

Enthalpy datum: CoolProp shifts h and s to its own reference state; that shift is a constant the EOS paper does not
fix.  `H_OFFSET` below is calibrated on the first golden row and verified on the other 200 (a constant offset).
"""
from __future__ import annotations

import numpy as np

from .helmholtz import HelmholtzFluid

_N = [0.924803575275, -0.492448489428, 0.661883336938, -1.92902649201, -0.0622469309629, 0.349943957581,
      0.564857472498, -1.61720005987, -0.481395031883, 0.421150636384, -0.0161962230825, 0.172100994165,
      0.00735448924933, 0.0168077305479, -0.00107626664179, -0.0137318088513, 0.000635466899859, 0.00304432279419,
      -0.0435762336045, -0.0723174889316, 0.0389644315272, -0.021220136391, 0.00408822981509, -0.0000551990017984,
      -0.0462016716479, -0.00300311716011, 0.0368825891208, -0.0025585684622, 0.00896915264558, -0.0044151337035,
      0.00133722924858, 0.000264832491957]
_D = [1, 1, 2, 2, 3, 3, 1, 1, 1, 3, 3, 4, 6, 6, 7, 7, 8, 8, 1, 2, 3, 4, 5, 8, 4, 5, 5, 8, 3, 5, 6, 9]
_T = [0.25, 0.875, 0.5, 0.875, 0.375, 0.75, 0.5, 0.75, 2, 1.25, 3.5, 1, 0.5, 3, 0, 2.75, 0.75, 2.5, 4, 6, 6, 3, 3, 6,
      16, 11, 15, 12, 12, 7, 4, 16]
_L = [0] * 6 + [1] * 12 + [2] * 6 + [3] * 5 + [4] * 3

NITROGEN = HelmholtzFluid(
    name="nitrogen", molar_mass=28.01348e-3, r_gas=8.31451, t_crit=126.192, rho_crit=11.1839e3, p_crit=3.3958e6,
    n=_N, d=_D, t=_T, l=_L,
    gn=[19.6688194015, -20.911560073, 0.0167788306989, 2627.67566274], gd=[1, 1, 3, 2], gt=[0, 1, 2, 3],
    geta=[20, 20, 15, 25], geps=[1, 1, 1, 1], gbeta=[325, 325, 300, 275], ggamma=[1.16, 1.16, 1.13, 1.25],
    a_lntau=2.5, a_const=-12.76952708, a_tau=-0.00784163,
    ideal_power=[(-1.934819e-4, -1.0), (-1.247742e-5, -2.0), (6.678326e-8, -3.0)],
    ideal_einstein=[(1.012941, 26.65788)],
)

# J/kg added to the EOS enthalpy (Span's datum) to land on CoolProp's; see the module docstring
H_OFFSET = 0.0


def viscosity(T, rho):
    """Pa s.  Lemmon & Jacobsen 2004, eqs. 1-4."""
    T, rho = np.asarray(T, dtype=np.float64), np.asarray(rho, dtype=np.float64)
    sigma, eps_k, M = 0.3656, 98.94, 28.01348
    b = (0.431, -0.4623, 0.08406, 0.005341, -0.00331)
    lt = np.log(T / eps_k)
    omega = np.exp(sum(bi * lt ** i for i, bi in enumerate(b)))
    eta0 = 0.0266958 * np.sqrt(M * T) / (sigma ** 2 * omega)                     # micro Pa s
    tau, delta = NITROGEN.t_crit / T, rho / NITROGEN.molar_mass / NITROGEN.rho_crit
    N = (10.72, 0.03989, 0.001208, -7.402, 4.620)
    t = (0.1, 0.25, 3.2, 0.9, 0.3)
    d = (2, 10, 12, 2, 1)
    l = (0, 1, 1, 2, 3)
    etar = sum(Ni * tau ** ti * delta ** di * (np.exp(-delta ** li) if li else 1.0) for Ni, ti, di, li in zip(N, t, d, l))
    return (eta0 + etar) * 1e-6, eta0


def conductivity(T, rho, cp, cv, dpdrho_T, mu):
    """W/m/K.  Lemmon & Jacobsen 2004, eqs. 5-11 (dilute + residual + simplified Olchowy-Sengers enhancement)."""
    T, rho = np.asarray(T, dtype=np.float64), np.asarray(rho, dtype=np.float64)
    fl = NITROGEN
    tau, delta = fl.t_crit / T, rho / fl.molar_mass / fl.rho_crit
    _, eta0 = viscosity(T, rho)
    lam0 = 1.511 * eta0 + 2.117 * tau ** -1.0 + -3.332 * tau ** -0.7              # mW/m/K
    N = (8.862, 31.11, -73.13, 20.03, -0.7096, 0.2672)
    t = (0.0, 0.03, 0.2, 0.8, 0.6, 1.9)
    d = (1, 2, 3, 4, 8, 10)
    l = (0, 0, 1, 2, 2, 2)
    lamr = sum(Ni * tau ** ti * delta ** di * (np.exp(-delta ** li) if li else 1.0) for Ni, ti, di, li in zip(N, t, d, l))
    # critical enhancement
    k_b, r0, nu, gam = 1.380658e-23, 1.01, 0.63, 1.2415
    qd, xi0, big_gamma, t_ref = 0.40e-9, 0.17e-9, 0.055, 252.384
    rhoc_mass = fl.rho_crit * fl.molar_mass

    def chi(Tv, drho_dp):
        return fl.p_crit * rho / rhoc_mass ** 2 * drho_dp

    drho_dp = 1.0 / dpdrho_T
    pr_ref = fl.properties(np.full(np.shape(rho), t_ref), rho)
    x = (chi(T, drho_dp) - chi(t_ref, 1.0 / pr_ref["dpdrho"]) * t_ref / T) / big_gamma
    xi = xi0 * np.where(x > 0, x, 0.0) ** (nu / gam)
    y = xi / qd
    with np.errstate(divide="ignore", invalid="ignore"):
        om = 2.0 / np.pi * ((cp - cv) / cp * np.arctan(y) + cv / cp * y)
        om0 = 2.0 / np.pi * (1.0 - np.exp(-1.0 / (1.0 / y + y ** 2 / 3.0 * (rhoc_mass / rho) ** 2)))
        lamc = rho * cp * k_b * r0 * T / (6.0 * np.pi * xi * mu) * (om - om0)
    lamc = np.where(xi > 0, lamc, 0.0)
    return (lam0 + lamr) * 1e-3 + lamc


def state(T, p):
    """The ten properties of the reference's alias table (simulation.py:89-100) + density, as a dict of arrays."""
    fl = NITROGEN
    rho = fl.density(T, p)
    pr = fl.properties(T, rho)
    mu, _ = viscosity(T, rho)
    k = conductivity(T, rho, pr["cp"], pr["cv"], pr["dpdrho"], mu)
    return {"beta": pr["beta"], "kappa": pr["kappa"], "prandtl": mu * pr["cp"] / k, "rho": rho, "mu": mu,
            "h": pr["h"] + H_OFFSET, "cp": pr["cp"], "cv": pr["cv"], "w": pr["w"], "k": k}
