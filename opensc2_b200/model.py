"""Operating-condition model of a conductor batch: what the reference evaluates on the
host before every `step()` (`Conductor.operating_conditions_em/_th`,
/root/reference/source_code/conductor.py:4591-4744; `eval_transp_coeff` :4822-5406;
`build_heat_source` :4423-4581), described as flat data for the device kernel K_P.

* `Model`   -- uniform across the batch (materials, correlations, HTC models): `osc2_model`.
* `Sweep`   -- swept per-conductor scalars (field, strain, heat slug, Tcs): `osc2_sweep`.
* `FluidTable` -- coolant properties on a regular (p, T) grid; the reference calls
  `CoolProp.PropsSI` (coolant.py:209-217), which is absent from this image: tables are
  an INPUT of the library (to be generated with CoolProp where it exists).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np

from ._abi import OSC2_N_FLUID_PROPS, Osc2Model, Osc2Sweep, c_dp

# fluid-table property order = the reference's alias table (simulation.py:89-100)
FP_BETA, FP_KAPPA, FP_PRANDTL, FP_RHO, FP_MU, FP_H, FP_CP, FP_CV, FP_SOUND, FP_K = range(10)
MAT_CU_NIST, MAT_NB3SN, MAT_SS, MAT_GE, MAT_AL, MAT_RE123, MAT_AG, MAT_HC276, MAT_SN60PB40, MAT_NBTI, MAT_MGB2 = range(11)
# SOLID_STACK: StackComponent (HTS tapes): materials = tape layers, weights = layer thicknesses, weight_sum = tape
# thickness (stack_component.py:262-330, 398-476)
SOLID_STRAND_MIXED, SOLID_STRAND_STAB, SOLID_JACKET, SOLID_STACK = range(4)


@dataclass(frozen=True)
class SolidModel:
    kind: int                               # SOLID_*
    materials: Tuple[int, ...]              # MAT_*, in the order of the reference's *_function arrays
    weights: Tuple[float, ...]              # homogenization_cefficients | cross_sections
    weight_sum: float                       # homogenization_cefficients_sum | inputs["CROSSECTION"]
    rrr: float = 100.0
    tc0m: float = 0.0
    bc20m: float = 0.0
    heated: bool = False                    # IQFUN == 1
    c0: float = 0.0                         # inputs["c0"], Jc scaling constant (A T / m^2); read by osc2_eval_tcs only


@dataclass(frozen=True)
class Model:
    ch_fluid: Tuple[int, ...]
    ch_ifriction: Tuple[int, ...]           # correlations.SUPPORTED_IFRICTION (-99, 110-115, 118, 121, 122, 124, 200-207, 209-211)
    ch_friction_mult: Tuple[float, ...]
    ch_htc_corr: Tuple[int, ...]            # 1 | 119 | 120 | 121 | 211 | 212
    solids: Tuple[SolidModel, ...]
    fs: Tuple[Tuple[int, float, float], ...] = ()                        # (HTC_choice, contact_HTC, multiplier)
    ff: Tuple[Tuple[int, float, float, float], ...] = ()                 # (+ interf_thickness)
    ss: Tuple[Tuple[int, float, float, float, float, float], ...] = ()   # (+ thick_rc, thick_cr, R_contact)
    es_htc: Tuple[float, ...] = ()          # constant convective HTC incl. Phi_conv and multiplier
    ch_roughness: Tuple[float, ...] = ()    # Channel.inputs["Roughness"], used by IFRICTION 121, 122
    ch_void_fraction: Tuple[float, ...] = ()   # Channel.inputs["VOID_FRACTION"], used by the bundle laws 204-206, 209
    ch_side_ratio: Tuple[float, ...] = ()      # Channel.inputs["SIDE2"] / ["SIDE1"], used by IFRICTION 119 (rectangular duct)
    ss_rad: Tuple[Tuple[float, float], ...] = ()   # per solid-solid interface: (denominator, multiplier) of choice 3

    def struct(self, topo) -> Osc2Model:
        assert len(self.ch_fluid) == topo.n_ch and len(self.solids) == topo.n_sol
        assert len(self.fs) == len(topo.fs) and len(self.ff) == len(topo.ff)
        assert len(self.ss) == len(topo.ss) and len(self.es_htc) == len(topo.es)
        m = Osc2Model()
        for j in range(topo.n_ch):
            m.ch_fluid[j], m.ch_ifriction[j] = self.ch_fluid[j], self.ch_ifriction[j]
            m.ch_friction_mult[j], m.ch_htc_corr[j] = self.ch_friction_mult[j], self.ch_htc_corr[j]
            m.ch_roughness[j] = self.ch_roughness[j] if self.ch_roughness else 0.0
            m.ch_void_fraction[j] = self.ch_void_fraction[j] if self.ch_void_fraction else 0.0
            m.ch_lam_coef[j] = duct_laminar_coefficient(self.ch_side_ratio[j]) if self.ch_side_ratio else 0.0
        for l, s in enumerate(self.solids):
            m.sol_kind[l], m.sol_nmat[l] = s.kind, len(s.materials)
            for i, (mat, w) in enumerate(zip(s.materials, s.weights)):
                m.sol_mat[l][i], m.sol_weight[l][i] = mat, w
            m.sol_weight_sum[l], m.sol_rrr[l] = s.weight_sum, s.rrr
            m.sol_tc0m[l], m.sol_bc20m[l], m.sol_heated[l] = s.tc0m, s.bc20m, int(s.heated)
            m.sol_c0[l] = s.c0
        for k, (c, h, mlt) in enumerate(self.fs):
            m.fs_choice[k], m.fs_htc[k], m.fs_mlt[k] = c, h, mlt
        for k, (c, h, mlt, th) in enumerate(self.ff):
            m.ff_choice[k], m.ff_htc[k], m.ff_mlt[k], m.ff_thick[k] = c, h, mlt, th
        for k, (c, h, mlt, t1, t2, rc) in enumerate(self.ss):
            m.ss_choice[k], m.ss_htc[k], m.ss_mlt[k] = c, h, mlt
            m.ss_thick_rc[k], m.ss_thick_cr[k], m.ss_rcontact[k] = t1, t2, rc
        for k, h in enumerate(self.es_htc):
            m.es_htc[k] = h
        for k in range(len(self.ss)):
            den, mlt = self.ss_rad[k] if self.ss_rad else (float("inf"), 1.0)
            m.ss_rad_den[k], m.ss_rad_mlt[k] = den, mlt
        return m


def duct_laminar_coefficient(alpha: float) -> float:
    """Numerator of the laminar friction law of a rectangular duct, rectangular_duct_demo_common_memo_laminar_flow_hole
    (channel.py:664-686), with the reference's own expression on Python floats (alpha = SIDE2 / SIDE1)."""
    return 24.0 * (1.0 - 1.3553 * alpha + 1.9467 * alpha ** 2 - 1.7012 * alpha ** 3 + 0.9564 * alpha ** 4 - 0.2537 * alpha ** 5)


@dataclass
class FluidTable:
    p_min: float
    p_step: float
    t_min: float
    t_step: float
    props: np.ndarray            # (10, n_p, n_t)

    def __post_init__(self):
        self.props = np.ascontiguousarray(self.props, dtype=np.float64)
        assert self.props.ndim == 3 and self.props.shape[0] == OSC2_N_FLUID_PROPS


@dataclass
class Sweep:
    """Batch-leading arrays; S = number of solids, E = elements."""
    b_field: np.ndarray          # (B, S, 2)  BISS, BOSS
    eps: np.ndarray              # (B, S)
    q0: np.ndarray               # (B, S)     W/m
    q_window: np.ndarray         # (B, S, 4)  first node, last node, TQBEG, TQEND
    tcs: Optional[np.ndarray] = None   # (B, S, E)
    jop: Optional[np.ndarray] = None   # (B, S) operating current density |I_op| / A_sc, A/m^2 (osc2_eval_tcs)

    def struct(self):
        keep = []
        s = Osc2Sweep()
        for name in ("b_field", "eps", "q0", "q_window", "tcs", "jop"):
            a = getattr(self, name)
            if a is None or a.size == 0:
                setattr(s, name, None)
                continue
            a = np.ascontiguousarray(a, dtype=np.float64)
            keep.append(a)
            setattr(s, name, a.ctypes.data_as(c_dp))
        return s, keep


def radiative_pair(emis_r: float, emis_c: float, outer_r: float, inner_r: float, outer_c: float, inner_c: float,
                   view_factor: float, multiplier: float) -> Tuple[float, float]:
    """(denominator, multiplier) of the radiative HTC between jackets r < c exactly as the reference resolves
    it (conductor.py:5161-5219 + _inner_radiative_htc :5428-5443): the inner jacket is the one whose outer
    perimeter is below the other's inner perimeter; the HTC multiplier is applied in the second branch only."""
    inf = float("inf")
    if not (emis_r > 0.0 and emis_c > 0.0 and view_factor > 0.0):
        return inf, 1.0
    vf_rec = float(np.reciprocal(view_factor))
    if outer_r < inner_c:
        return (1.0 - emis_r) / emis_r + vf_rec + (1.0 - emis_c) / emis_c * (outer_r / inner_c), 1.0
    if outer_c < inner_r:
        return (1.0 - emis_c) / emis_c + vf_rec + (1.0 - emis_r) / emis_r * (outer_c / inner_r), multiplier
    return inf, 1.0


def heat_window(zcoord: np.ndarray, xq_beg: float, xq_end: float) -> Tuple[int, int]:
    """`SolidComponent.heat0_where` (solid_component.py:523-531): inclusive node window."""
    lo = int(np.min(np.nonzero(zcoord >= xq_beg)))
    hi = int(np.max(np.nonzero(zcoord <= xq_end)))
    return lo, hi


# ----------------------------------------------------------------------------------------
# Shipped-deck shaped models
# ----------------------------------------------------------------------------------------

def case1_model(correlations: bool = False) -> Model:
    """CASE_1 ITER-like LTS: 2 He channels (f = 0.02), STR_MIX_1 (Cu RRR 197.71 + Nb3Sn,
    Cu:nonCu 2.0846), Z_JACKET_1 (SS 3.0699e-4 + glass-epoxy 2.7966e-4 m^2); the deck's
    interfaces are all constant (HTC_choice -2 / -1: 1000, 1000, 1000, 500, 0 W/m^2K).
    `correlations=True` switches the interfaces to the correlation-driven models
    (choice 2 / 1) to exercise them."""
    stab = 2.0846
    solids = (
        SolidModel(SOLID_STRAND_MIXED, (MAT_CU_NIST, MAT_NB3SN), (stab, 1.0), float(np.array([stab, 1.0]).sum()),
                   rrr=197.71, tc0m=16.22, bc20m=32.35, heated=True, c0=1.21508e11),
        SolidModel(SOLID_JACKET, (MAT_SS, MAT_GE), (3.0699e-4, 2.7966e-4), 3.0699e-4 + 2.7966e-4),
    )
    if correlations:
        fs = ((2, 0.0, 1.0), (2, 0.0, 0.8))
        ff = ((2, 0.0, 1.0, 1.0e-3),)
        ss = ((1, 0.0, 1.0, 1.0e-3, 2.0e-3, 5.0e-4),)
    else:
        fs = ((-2, 1000.0, 1.0), (-2, 1000.0, 1.0))
        ff = ((-2, 1000.0, 1.0, 1.0e-3),)
        ss = ((-1, 500.0, 1.0, 0.0, 0.0, 0.0),)
    return Model(ch_fluid=(0, 0), ch_ifriction=(-99, 124 if correlations else -99), ch_friction_mult=(0.02, 0.02),
                 ch_htc_corr=(1, 212 if correlations else 1), solids=solids, fs=fs, ff=ff, ss=ss, es_htc=(0.0,))


def case2_mini_model() -> Model:
    """Model of `problem.case2_mini_topology`: IFRICTION 122 / 124, Nusselt 1 / 212, Kapitza-limited
    channel-solid HTCs (choice 2), one conductive and one constant solid-solid contact."""
    solids = (
        SolidModel(SOLID_STRAND_STAB, (MAT_AL,), (1.0,), 1.0, heated=True),
        SolidModel(SOLID_STRAND_STAB, (MAT_CU_NIST,), (1.0,), 1.0, rrr=100.0),
        SolidModel(SOLID_JACKET, (MAT_SS,), (1.3972e-4,), 1.3972e-4),
    )
    return Model(ch_fluid=(0, 0), ch_ifriction=(122, 124), ch_friction_mult=(0.02, 0.02), ch_htc_corr=(1, 212),
                 ch_roughness=(2.0e-6, 0.0), solids=solids,
                 fs=((2, 0.0, 1.0), (2, 0.0, 1.0), (2, 0.0, 1.0), (-2, 0.0, 1.0)),
                 ss=((-1, 2.0e5, 1.0, 0.0, 0.0, 0.0), (1, 0.0, 1.0, 5.0e-4, 1.0e-3, 2.0e-4)), es_htc=(0.0,))


def case2_model(topo) -> Model:
    """CASE_2 ENEA HTS CICC on `problem.case2_topology`: 12 slot channels
    (IFRICTION 124, Nusselt 212) + the central channel (122, Nusselt 1); 6 tape stacks restated as
    single-material RE123 solids (the legacy STR_SC class is absent from HEAD, Appendix B), 18 aluminium
    stabilisers, one SS jacket; channel-solid HTCs by correlation (choice 2) except channel-jacket
    (-2, 0 W/m^2K); solid-solid contacts constant (-1)."""
    M, nsc, nst = 13, 6, 18
    S = nsc + nst + 1
    jk = S - 1
    solids = tuple([SolidModel(SOLID_STRAND_STAB, (MAT_RE123,), (1.0,), 1.0, heated=(l == 0)) for l in range(nsc)]
                   + [SolidModel(SOLID_STRAND_STAB, (MAT_AL,), (1.0,), 1.0) for _ in range(nst)]
                   + [SolidModel(SOLID_JACKET, (MAT_SS,), (1.3972e-4,), 1.3972e-4)])
    fs = tuple((-2, 0.0, 1.0) if l == jk else (2, 0.0, 1.0) for (_j, l) in topo.fs)
    def ss_const(a, b):
        if b == jk:
            return 0.0 if a < nsc else 1140.0
        return 32750.0 if a < nsc else 2.0e5
    ss = tuple((-1, ss_const(a, b), 1.0, 0.0, 0.0, 0.0) for (a, b) in topo.ss)
    return Model(ch_fluid=(0,) * M, ch_ifriction=(124,) * 12 + (122,), ch_friction_mult=(0.02,) * M,
                 ch_htc_corr=(212,) * 12 + (1,), ch_roughness=(0.0,) * 12 + (2.0e-6,), solids=solids,
                 fs=fs, ff=(), ss=ss, es_htc=(0.0,))


CASE3_MINI_JACKETS = dict(emissivity=(0.8, 0.8), outer_perimeter=(0.060, 0.110), inner_perimeter=(0.050, 0.090),
                          view_factor=1.0)


def case3_mini_model() -> Model:
    """Model of `problem.case3_mini_topology`: Kapitza-limited channel-solid HTCs, radiation between jackets 1
    and 2 (HTC_choice 3, emissivity 0.8, view factor 1), a constant contact (-1, 500 W/m^2K) between jackets
    2 and 3, constant convection to the environment (-4, 5 W/m^2K)."""
    solids = (
        SolidModel(SOLID_STRAND_STAB, (MAT_CU_NIST,), (1.0,), 1.0, rrr=100.0),
        SolidModel(SOLID_JACKET, (MAT_SS,), (2.0e-4,), 2.0e-4),
        SolidModel(SOLID_JACKET, (MAT_SS,), (1.5e-4,), 1.5e-4),
        SolidModel(SOLID_JACKET, (MAT_GE,), (3.0e-4,), 3.0e-4),
    )
    j = CASE3_MINI_JACKETS
    rad = radiative_pair(j["emissivity"][0], j["emissivity"][1], j["outer_perimeter"][0], j["inner_perimeter"][0],
                         j["outer_perimeter"][1], j["inner_perimeter"][1], j["view_factor"], 1.0)
    return Model(ch_fluid=(0,), ch_ifriction=(-99,), ch_friction_mult=(0.02,), ch_htc_corr=(1,), solids=solids,
                 fs=((2, 0.0, 1.0), (2, 0.0, 1.0)), ss=((3, 0.0, 1.0, 0.0, 0.0, 0.0), (-1, 500.0, 1.0, 0.0, 0.0, 0.0)),
                 ss_rad=(rad, (float("inf"), 1.0)), es_htc=(5.0,))


def synthetic_helium_table(n_p: int = 65, n_t: int = 129) -> FluidTable:
    """Smooth stand-in for supercritical helium near 4.5 K / 6 bar (CoolProp is absent here;
    magnitudes follow SURVEY.md Appendix A).  NOT physical data: benchmarks and parity tests
    only need the same table on both sides."""
    p = np.linspace(4.0e5, 8.0e5, n_p)
    t = np.linspace(3.5, 12.0, n_t)
    P, T = np.meshgrid(p, t, indexing="ij")
    dp, dt = (P - 6.0e5) / 1.0e5, T - 4.5
    rho = 140.0 - 8.0 * dt + 0.35 * dt ** 2 + 2.0 * dp
    cv = 2.5e3 + 200.0 * dt - 6.0 * dt ** 2 + 30.0 * dp
    cp = 3.9e3 + 500.0 * dt - 25.0 * dt ** 2 - 60.0 * dp
    props = np.empty((OSC2_N_FLUID_PROPS, n_p, n_t))
    props[FP_BETA] = 0.06 + 0.02 * dt - 0.002 * dp
    props[FP_KAPPA] = 2.0e-7 * (1.0 + 0.15 * dt - 0.05 * dp)
    props[FP_PRANDTL] = 0.75 + 0.03 * dt
    props[FP_RHO] = rho
    props[FP_MU] = 3.5e-6 * (1.0 - 0.03 * dt + 0.02 * dp)
    props[FP_H] = 1.2e4 + 4.0e3 * dt + 300.0 * dp
    props[FP_CP] = cp
    props[FP_CV] = cv
    props[FP_SOUND] = 220.0 + 15.0 * dt + 8.0 * dp
    props[FP_K] = 0.02 * (1.0 + 0.02 * dt)
    return FluidTable(float(p[0]), float(p[1] - p[0]), float(t[0]), float(t[1] - t[0]), props)


def helium_gas_table(n_p: int = 65, n_t: int = 161, p_range=(0.5e5, 4.0e5), t_range=(20.0, 60.0)) -> FluidTable:
    """Stand-in for gaseous helium around 30 K / 1-2.5 bar (CASE_3 CHAN_1): ideal gas with a constant second virial
    coefficient, cp = 5/2 R, power-law viscosity, constant Prandtl number, anchored on the values the reference's
    golden CASE_3 CHAN_1.tsv prints at 30 K (mu 4.675e-6 Pa s, Pr 0.723, h 160 370 J/kg).  NOT CoolProp: within about
    1 % of real helium here, which is all a solver test needs -- both sides of every parity test read the same table."""
    rs = 8.3144598 / 4.002602e-3
    p = np.linspace(p_range[0], p_range[1], n_p)
    t = np.linspace(t_range[0], t_range[1], n_t)
    P, T = np.meshgrid(p, t, indexing="ij")
    bvir = 2.4e-6 / 4.002602e-3                   # m^3/kg, B(30 K) ~ +2.4 cm^3/mol
    rho = P / (rs * T + bvir * P)                 # v = R T / p + B
    cp, cv = 2.5 * rs * np.ones_like(T), 1.5 * rs * np.ones_like(T)
    mu = 4.675e-6 * (T / 30.0) ** 0.6286
    props = np.empty((OSC2_N_FLUID_PROPS, n_p, n_t))
    props[FP_BETA] = (rs / P) * rho               # (1/v) (dv/dT)_p
    props[FP_KAPPA] = (rs * T / P ** 2) * rho     # -(1/v) (dv/dp)_T
    props[FP_PRANDTL] = 0.723
    props[FP_RHO] = rho
    props[FP_MU] = mu
    props[FP_H] = cp * T + bvir * P + 4580.0
    props[FP_CP] = cp
    props[FP_CV] = cv
    props[FP_SOUND] = np.sqrt(cp / cv * rs * T) * (1.0 + bvir * rho)
    props[FP_K] = mu * cp / 0.723
    return FluidTable(float(p[0]), float(p[1] - p[0]), float(t[0]), float(t[1] - t[0]), props)


def synthetic_sweep(topo, model: Model, n_nodes: int, batch: int, seed: int = 20261019, length: float = 10.0,
                    with_tcs: bool = True) -> Sweep:
    """Swept scalars of SURVEY.md 8(d): Q0 ~ U[50, 1000] W/m on x in [1, 3] m for t in [10, 20] s,
    background field ~ U[0, 12] T.  Tcs here is a smooth synthetic profile (the real one is the
    reference's bisection, evaluated once at step 0 on the host)."""
    rng = np.random.default_rng(seed)
    S, E = topo.n_sol, n_nodes - 1
    z = np.linspace(0.0, length, n_nodes)
    lo, hi = heat_window(z, 0.1 * length, 0.3 * length)
    b_field = np.zeros((batch, S, 2))
    b0 = rng.uniform(0.0, 12.0, (batch, 1))
    b_field[:, :, 0] = b0
    b_field[:, :, 1] = b0 * rng.uniform(0.9, 1.1, (batch, 1))
    eps = np.tile(rng.uniform(-0.006, 0.0, (batch, 1)), (1, S))
    q0 = rng.uniform(50.0, 1000.0, (batch, S))
    q_window = np.empty((batch, S, 4))
    q_window[..., 0], q_window[..., 1], q_window[..., 2], q_window[..., 3] = lo, hi, 10.0, 20.0
    tcs = None
    if with_tcs:
        tcs = 6.0 + 2.0 * rng.uniform(0.0, 1.0, (batch, S, 1)) + 0.5 * np.sin(np.linspace(0.0, 3.0, E))[None, None, :]
    return Sweep(b_field=b_field, eps=eps, q0=q0, q_window=q_window, tcs=tcs)


def table_lookup(tab: FluidTable, prop: int, p, t):
    """Host-side bilinear lookup with the arithmetic of the device kernel (initial states only)."""
    p, t = np.asarray(p, dtype=np.float64), np.asarray(t, dtype=np.float64)
    n_p, n_t = tab.props.shape[1:]
    fp = np.clip(np.floor((p - tab.p_min) / tab.p_step), 0, n_p - 2)
    ft = np.clip(np.floor((t - tab.t_min) / tab.t_step), 0, n_t - 2)
    a = (p - (tab.p_min + fp * tab.p_step)) / tab.p_step
    b = (t - (tab.t_min + ft * tab.t_step)) / tab.t_step
    ip, it = fp.astype(int), ft.astype(int)
    f = tab.props[prop]
    lo = f[ip, it] + b * (f[ip, it + 1] - f[ip, it])
    hi = f[ip + 1, it] + b * (f[ip + 1, it + 1] - f[ip + 1, it])
    return lo + a * (hi - lo)


def consistent_initial_state(topo, model: Model, tab: FluidTable, n_nodes: int, batch: int, seed: int = 20261019,
                             length: float = 10.0, dt: float = 0.1, p_in: float = 6.0e5, p_out: float = 5.9e5):
    """A physically consistent start of a CASE_1-like transient, in the spirit of the reference's
    initialisation (coolant.py:87-130, solid_components_initialization.py:24-52): linear pressure,
    uniform temperature (swept inlet T ~ U[4.3, 4.7] K), velocities at the friction/pressure-drop
    balance, solids at the coolant temperature.  Returns a StepArrays whose geometry, boundary values
    and state are filled; the coefficient arrays are placeholders K_P overwrites."""
    from .problem import (BC_MDTIN, BC_MDTOUT, BC_PINL, BC_POUT, BC_TINL, BC_TOUT, synthetic_step_arrays)

    arr = synthetic_step_arrays(topo, n_nodes, batch=batch, seed=seed, length=length, dt=dt, heated=False)
    rng = np.random.default_rng(seed + 1)
    M, S, d, N = topo.n_ch, topo.n_sol, topo.ndf, n_nodes
    xi = np.linspace(0.0, 1.0, N)
    t_in = rng.uniform(4.3, 4.7, (batch, 1))
    sv = np.empty((batch, N, d))
    sgn = np.where(np.array(topo.ch_forward), 1.0, -1.0)
    for j in range(M):
        pj = (p_in + (p_out - p_in) * xi)[None, :] if sgn[j] > 0 else (p_out + (p_in - p_out) * xi)[None, :]
        pj = np.broadcast_to(pj, (batch, N))
        tj = np.broadcast_to(t_in, (batch, N))
        rho = table_lookup(tab, FP_RHO, pj, tj)
        f = model.ch_friction_mult[j]                      # IFRICTION -99: f = 1 * multiplier
        v = sgn[j] * np.sqrt(abs(p_in - p_out) / length * topo.ch_dh[j] / (2.0 * f * rho))
        sv[:, :, j], sv[:, :, M + j], sv[:, :, 2 * M + j] = v, pj, tj
        arr.fl_bc[:, BC_PINL, j], arr.fl_bc[:, BC_POUT, j] = p_in, p_out
        arr.fl_bc[:, BC_TINL, j], arr.fl_bc[:, BC_TOUT, j] = t_in[:, 0], t_in[:, 0]
        mdot = topo.ch_area[j] * v[:, 0] * rho[:, 0]
        arr.fl_bc[:, BC_MDTIN, j], arr.fl_bc[:, BC_MDTOUT, j] = mdot, mdot
    for l in range(S):
        sv[:, :, 3 * M + l] = t_in
    arr.sysvar = sv.reshape(batch, N * d)
    arr.syslod = np.zeros((batch, N * d, 2))
    arr.qsource = np.zeros((batch, N, S))
    arr.num_step = 1
    return arr
