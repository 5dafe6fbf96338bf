"""Generic multiparameter Helmholtz-energy equation of state, vectorised over numpy arrays.

    a(T, rho) / (R T) = alpha0(tau, delta) + alphar(tau, delta),   tau = Tc / T, delta = rho / rhoc

alphar = sum n_i delta^d_i tau^t_i [exp(-delta^l_i)]                                  (polynomial / exponential terms)
       + sum n_i delta^d_i tau^t_i exp(-eta_i (delta - eps_i)^2 - beta_i (tau - gamma_i)^2)          (Gaussian terms)
alpha0 = ln delta + a1 ln tau + a2 + a3 tau + sum_k c_k tau^(-k) + sum_j v_j ln(1 - exp(-u_j tau))

All thermodynamic properties follow from alpha and its first and second derivatives by the standard relations
(Span, "Multiparameter Equations of State", 2000, Table 3.10); they are what `CoolProp.PropsSI` evaluates for the
reference (/root/reference/source_code/coolant.py:202-222, alias table simulation.py:89-100):
    p = rho R T (1 + delta ar_d)                          cv = -R tau^2 (a0_tt + ar_tt)
    cp = cv + R (1 + delta ar_d - delta tau ar_dt)^2 / (1 + 2 delta ar_d + delta^2 ar_dd)
    w^2 = R T [1 + 2 delta ar_d + delta^2 ar_dd - (1 + delta ar_d - delta tau ar_dt)^2 / (tau^2 (a0_tt + ar_tt))]
    h = R T [1 + tau (a0_t + ar_t) + delta ar_d]
    (dp/drho)_T = R T (1 + 2 delta ar_d + delta^2 ar_dd)   (dp/dT)_rho = rho R (1 + delta ar_d - delta tau ar_dt)
    kappa_T = 1 / (rho (dp/drho)_T)                        beta = (dp/dT)_rho / (rho (dp/drho)_T)
with R the specific gas constant.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Sequence

import numpy as np


@dataclass
class HelmholtzFluid:
    name: str
    molar_mass: float            # kg/mol
    r_gas: float                 # J/mol/K (the value the EOS was fitted with)
    t_crit: float                # K (reducing temperature)
    rho_crit: float              # mol/m^3 (reducing density)
    p_crit: float                # Pa
    # residual part
    n: Sequence[float] = ()
    d: Sequence[float] = ()
    t: Sequence[float] = ()
    l: Sequence[float] = ()      # 0 -> polynomial term
    gn: Sequence[float] = ()     # Gaussian terms
    gd: Sequence[float] = ()
    gt: Sequence[float] = ()
    geta: Sequence[float] = ()
    geps: Sequence[float] = ()
    gbeta: Sequence[float] = ()
    ggamma: Sequence[float] = ()
    # ideal part: ln(delta) + a_lntau ln(tau) + a_const + a_tau tau + sum c_k tau^k_pow + sum v ln(1 - exp(-u tau))
    a_lntau: float = 0.0
    a_const: float = 0.0
    a_tau: float = 0.0
    ideal_power: Sequence[Sequence[float]] = ()    # (coefficient, exponent of tau)
    ideal_einstein: Sequence[Sequence[float]] = ()  # (v, u)

    @property
    def rs(self) -> float:
        return self.r_gas / self.molar_mass

    # ---- Helmholtz derivatives ------------------------------------------------------------------------------
    def alphar(self, delta, tau) -> Dict[str, np.ndarray]:
        delta, tau = np.asarray(delta, dtype=np.float64), np.asarray(tau, dtype=np.float64)
        out = {k: np.zeros(np.broadcast(delta, tau).shape) for k in ("a", "d", "dd", "t", "tt", "dt")}
        for n, d, t, l in zip(self.n, self.d, self.t, self.l):
            if l:
                u = delta ** l
                A = n * delta ** d * tau ** t * np.exp(-u)
                g1 = d - l * u
                g2 = g1 * (g1 - 1.0) - l * l * u
            else:
                A = n * delta ** d * tau ** t
                g1 = d
                g2 = d * (d - 1.0)
            out["a"] += A
            out["d"] += A * g1 / delta
            out["dd"] += A * g2 / delta ** 2
            out["t"] += A * t / tau
            out["tt"] += A * t * (t - 1.0) / tau ** 2
            out["dt"] += A * g1 * t / (delta * tau)
        for n, d, t, eta, eps, beta, gam in zip(self.gn, self.gd, self.gt, self.geta, self.geps, self.gbeta, self.ggamma):
            A = n * delta ** d * tau ** t * np.exp(-eta * (delta - eps) ** 2 - beta * (tau - gam) ** 2)
            fd = d / delta - 2.0 * eta * (delta - eps)
            ft = t / tau - 2.0 * beta * (tau - gam)
            out["a"] += A
            out["d"] += A * fd
            out["dd"] += A * (fd * fd - d / delta ** 2 - 2.0 * eta)
            out["t"] += A * ft
            out["tt"] += A * (ft * ft - t / tau ** 2 - 2.0 * beta)
            out["dt"] += A * fd * ft
        return out

    def alpha0(self, delta, tau) -> Dict[str, np.ndarray]:
        tau = np.asarray(tau, dtype=np.float64)
        a = np.log(delta) + self.a_lntau * np.log(tau) + self.a_const + self.a_tau * tau
        at = self.a_lntau / tau + self.a_tau
        att = -self.a_lntau / tau ** 2
        for c, k in self.ideal_power:
            a = a + c * tau ** k
            at = at + c * k * tau ** (k - 1.0)
            att = att + c * k * (k - 1.0) * tau ** (k - 2.0)
        for v, u in self.ideal_einstein:
            ex = np.exp(-u * tau)
            a = a + v * np.log(1.0 - ex)
            at = at + v * u * ex / (1.0 - ex)
            att = att - v * u * u * ex / (1.0 - ex) ** 2
        return {"a": a, "t": at, "tt": att}

    # ---- state functions of (T, rho_mass) -------------------------------------------------------------------
    def pressure(self, T, rho):
        delta, tau = rho / self.molar_mass / self.rho_crit, self.t_crit / T
        return rho * self.rs * T * (1.0 + delta * self.alphar(delta, tau)["d"])

    def properties(self, T, rho) -> Dict[str, np.ndarray]:
        T, rho = np.asarray(T, dtype=np.float64), np.asarray(rho, dtype=np.float64)
        delta, tau = rho / self.molar_mass / self.rho_crit, self.t_crit / T
        r, z = self.alphar(delta, tau), self.alpha0(delta, tau)
        rs = self.rs
        a1 = 1.0 + delta * r["d"] - delta * tau * r["dt"]
        a2 = 1.0 + 2.0 * delta * r["d"] + delta ** 2 * r["dd"]
        ctt = tau ** 2 * (z["tt"] + r["tt"])
        cv = -rs * ctt
        cp = cv + rs * a1 ** 2 / a2
        dpdrho = rs * T * a2
        dpdT = rho * rs * a1
        return {"p": rho * rs * T * (1.0 + delta * r["d"]), "cv": cv, "cp": cp,
                "w": np.sqrt(rs * T * (a2 - a1 ** 2 / ctt)),
                "h": rs * T * (1.0 + tau * (z["t"] + r["t"]) + delta * r["d"]),
                "kappa": 1.0 / (rho * dpdrho), "beta": dpdT / (rho * dpdrho), "dpdrho": dpdrho, "dpdT": dpdT}

    def density(self, T, p, rho_guess=None, tol: float = 1e-13, max_iter: int = 100):
        """rho(T, p) by Newton iteration on p(T, rho) from the ideal-gas density (single-phase gas and supercritical
        states; a liquid state needs `rho_guess` on the liquid side)."""
        T, p = np.broadcast_arrays(np.asarray(T, dtype=np.float64), np.asarray(p, dtype=np.float64))
        rho = p / (self.rs * T) if rho_guess is None else np.array(np.broadcast_to(rho_guess, T.shape), dtype=np.float64)
        rho = np.array(rho, dtype=np.float64, copy=True)
        for _ in range(max_iter):
            delta, tau = rho / self.molar_mass / self.rho_crit, self.t_crit / T
            r = self.alphar(delta, tau)
            f = rho * self.rs * T * (1.0 + delta * r["d"]) - p
            df = self.rs * T * (1.0 + 2.0 * delta * r["d"] + delta ** 2 * r["dd"])
            step = f / df
            step = np.clip(step, -0.5 * rho, 0.5 * rho)        # damped: stay on the starting branch
            rho = rho - step
            if np.all(np.abs(step) <= tol * rho):
                break
        return rho
