"""Neutral description of one OPENSC2 `step()` problem (topology + state).

The reference keeps everything `step()` reads scattered over the `Conductor`
object graph (SURVEY.md section 8b lists the attributes,
`/root/reference/source_code/utility_functions/transient_solution_functions.py:186-599`
is the consumer).  `Topology` + `StepArrays` hold exactly that information as
flat numpy arrays so that the same problem can be handed to

* the reference itself (tests/golden/ref_harness.py builds a duck-typed
  conductor from it -- this container only),
* the CPU oracle (oracle/oracle.py),
* the batched CUDA library (opensc2_b200.batch).

Equation order inside a node follows `Conductor.__build_equation_idx`
(`/root/reference/source_code/conductor.py:897-946`):
``v_1..v_M, p_1..p_M, T_1..T_M, Ts_1..Ts_S``; global index = node*d + eq.
All per-conductor arrays carry an optional leading batch axis ``B``.
"""
from __future__ import annotations

from dataclasses import dataclass, field, fields, replace
from typing import List, Tuple

import numpy as np

# Row order of the per-channel Gauss-point coefficient block `fl_gauss`.
FL_V, FL_P, FL_T, FL_RHO, FL_C0, FL_GRUN, FL_H, FL_CV, FL_F = range(9)
N_FL_GAUSS = 9
# Row order of the per-solid Gauss-point coefficient block `sol_gauss`.
SOL_RHO, SOL_CP, SOL_K = range(3)
N_SOL_GAUSS = 3
# Row order of `fl_bc` (boundary-condition values per channel).
BC_PINL, BC_POUT, BC_TINL, BC_TOUT, BC_MDTIN, BC_MDTOUT = range(6)
N_BC = 6
# Row order of `fl_node` (nodal values the boundary conditions look at).
ND_VFIRST, ND_VLAST, ND_RHOFIRST, ND_RHOLAST = range(4)
N_ND = 4


@dataclass(frozen=True)
class Topology:
    """Everything that is uniform across a batch of conductor variants."""

    n_ch: int
    n_sol: int
    ch_area: Tuple[float, ...]            # channel.inputs["CROSSECTION"]
    ch_dh: Tuple[float, ...]              # channel.inputs["HYDIAMETER"]
    ch_forward: Tuple[bool, ...]          # channel.flow_dir[0] == "forward"
    ch_intial: Tuple[int, ...]            # abs(coolant.operations["INTIAL"]) in {1,2,3}
    sol_area: Tuple[float, ...]           # solid.inputs["CROSSECTION"]
    sol_costeta: Tuple[float, ...]        # solid.inputs["COSTETA"]
    ff: Tuple[Tuple[int, int], ...] = ()  # fluid-fluid interfaces (channel idx pairs)
    fs: Tuple[Tuple[int, int], ...] = ()  # fluid-solid interfaces (channel idx, solid idx)
    ss: Tuple[Tuple[int, int], ...] = ()  # solid-solid interfaces (solid idx pairs)
    es: Tuple[int, ...] = ()              # environment-solid interfaces (solid idx)
    es_choice2: Tuple[bool, ...] = ()     # HTC_choice.at[env, jacket] == 2 for that interface
    is_rectangular: bool = False          # conductor.inputs["Is_rectangular"]
    height: float = 0.0                   # conductor.inputs["Height"]
    width: float = 0.0                    # conductor.inputs["Width"]
    delta_p_min: float = 1.0e-4           # Conductor.Delta_p_min (conductor.py:89-94)
    k_loc: float = 1.0                    # Conductor.k_loc
    lambda_v: float = 1.0                 # Conductor.lambda_v
    theta: float = 1.0                    # conductor.theta_method
    method: str = "BE"                    # conductor.inputs["METHOD"] in {"BE","CN"}

    @property
    def ndf(self) -> int:
        return 3 * self.n_ch + self.n_sol

    def eq_v(self, j: int) -> int:
        return j

    def eq_p(self, j: int) -> int:
        return self.n_ch + j

    def eq_t(self, j: int) -> int:
        return 2 * self.n_ch + j

    def eq_s(self, l: int) -> int:
        return 3 * self.n_ch + l

    def validate(self) -> None:
        M, S = self.n_ch, self.n_sol
        assert len(self.ch_area) == len(self.ch_dh) == len(self.ch_forward) == len(self.ch_intial) == M
        assert len(self.sol_area) == len(self.sol_costeta) == S
        assert all(i in (1, 2, 3) for i in self.ch_intial)
        assert all(0 <= a < M and 0 <= b < M and a != b for a, b in self.ff)
        assert all(0 <= a < M and 0 <= b < S for a, b in self.fs)
        assert all(0 <= a < S and 0 <= b < S and a != b for a, b in self.ss)
        assert all(0 <= a < S for a in self.es)
        assert len(self.es_choice2) == len(self.es)
        assert self.method in ("BE", "CN")


@dataclass
class StepArrays:
    """Per-conductor inputs/outputs of one `step()`; leading batch axis optional.

    Shapes below omit the batch axis.  E = N-1 elements, Total = d*N.
    """

    dz: np.ndarray            # (E,)  grid_features["delta_z"]
    fl_gauss: np.ndarray      # (9, M, E)  v,p,T,rho,c0,Gruneisen,h,cv,f_total at Gauss points
    fl_node: np.ndarray       # (4, M)  nodal v[0], v[-1], rho[0], rho[-1] (old state)
    fl_bc: np.ndarray         # (6, M)  PREINL, PREOUT, TEMINL, TEMOUT, MDTIN, MDTOUT
    sol_gauss: np.ndarray     # (3, S, E)  rho, cp, k at Gauss points
    sol_q: np.ndarray         # (2, S, E, 2)  Q1,Q2 columns [present, previous]
    peri_ff: np.ndarray       # (2, nff, E)  open, close contact perimeter
    htc_ff: np.ndarray        # (2, nff, E)  open, close heat transfer coefficient
    peri_fs: np.ndarray       # (nfs, E)
    htc_fs: np.ndarray        # (nfs, E)
    peri_ss: np.ndarray       # (nss, E)
    htc_ss: np.ndarray        # (nss, E)   HTC["sol_sol"]["cond"]
    peri_es: np.ndarray       # (nes, E)
    htc_es: np.ndarray        # (nes, 3, E)  conv (or side), bottom, top
    qsource: np.ndarray       # (N, S)
    sysvar: np.ndarray        # (Total,)  dict_Step["SYSVAR"][:,0]
    syslod: np.ndarray        # (Total, 2)  dict_Step["SYSLOD"]
    t_env: float = 300.0      # environment.inputs["Temperature"]
    dt: float = 0.1           # conductor.time_step
    num_step: int = 2         # conductor.cond_num_step
    peri_ss_nodal: np.ndarray = None   # (nss, N) nodal contact perimeters, radiative interfaces only (optional)

    def copy(self) -> "StepArrays":
        kw = {}
        for f in fields(self):
            v = getattr(self, f.name)
            kw[f.name] = v.copy() if isinstance(v, np.ndarray) else v
        return StepArrays(**kw)

    @property
    def n_nodes(self) -> int:
        return self.dz.shape[-1] + 1


# ----------------------------------------------------------------------------
# Shipped-deck shaped topologies (Appendix A of SURVEY.md).
# ----------------------------------------------------------------------------

def case1_topology(intial: int = 1, **kw) -> Topology:
    """CASE_1 ITER-like LTS: 2 He channels, STR_MIX_1, Z_JACKET_1; d = 8."""
    t = Topology(
        n_ch=2, n_sol=2,
        ch_area=(5.0265e-5, 3.6965e-4), ch_dh=(8e-3, 3.2676e-4),
        ch_forward=(True, True), ch_intial=(intial, intial),
        sol_area=(0.000753951765, 3.0699e-4 + 2.7966e-4), sol_costeta=(0.9699, 1.0),
        ff=((0, 1),), fs=((1, 0), (1, 1)), ss=((0, 1),), es=(1,), es_choice2=(False,),
    )
    return replace(t, **kw)


def case3_topology(**kw) -> Topology:
    """CASE_3 HTS HVDC: He forward + N2 backward, STR_STAB_1, 6 jackets; d = 13."""
    t = Topology(
        n_ch=2, n_sol=7,
        ch_area=(2.356e-4, 3.393e-3), ch_dh=(0.015, 0.04408),
        ch_forward=(True, False), ch_intial=(3, 3),
        sol_area=(1.0e-4, 2.0e-4, 1.5e-4, 3.0e-4, 3.2e-4, 4.0e-4, 5.0e-4),
        sol_costeta=(1.0,) * 7,
        ff=(), fs=((0, 0), (0, 1), (1, 3), (1, 4)),
        ss=((1, 2), (2, 3), (4, 5), (5, 6)), es=(6,), es_choice2=(False,),
    )
    return replace(t, **kw)


def case2_mini_topology(**kw) -> Topology:
    """A small CASE_2-flavoured conductor (d = 9, generic kernels): two He channels with correlation
    driven friction (122, 124) and Nusselt (1, 212), an aluminium and a copper stabiliser, a
    stainless-steel jacket; no channel-channel interface (like CASE_2)."""
    t = Topology(
        n_ch=2, n_sol=3,
        ch_area=(1.963e-5, 6.5e-7), ch_dh=(2.499e-3, 2.9115e-4),
        ch_forward=(True, True), ch_intial=(3, 3),
        sol_area=(2.10568e-5, 1.13187e-5, 1.3972e-4), sol_costeta=(1.0, 1.0, 1.0),
        ff=(), fs=((0, 0), (0, 1), (1, 1), (1, 2)), ss=((0, 1), (1, 2)), es=(2,), es_choice2=(False,),
    )
    return replace(t, **kw)


def case3_mini_topology(**kw) -> Topology:
    """A small CASE_3-flavoured conductor (d = 7, generic kernels): one He channel cooling a copper
    stabiliser and the innermost jacket, a radiative gap between jackets 1 and 2, a contact between
    jackets 2 and 3, convection to the environment on the outermost one."""
    t = Topology(
        n_ch=1, n_sol=4,
        ch_area=(2.356e-4,), ch_dh=(0.015,), ch_forward=(True,), ch_intial=(3,),
        sol_area=(1.0e-4, 2.0e-4, 1.5e-4, 3.0e-4), sol_costeta=(1.0,) * 4,
        ff=(), fs=((0, 0), (0, 1)), ss=((1, 2), (2, 3)), es=(3,), es_choice2=(False,),
    )
    return replace(t, **kw)


def case2_topology(**kw) -> Topology:
    """CASE_2 ENEA HTS CICC: 13 channels, 6 SC + 18 stabiliser + 1 jacket; d = 64.

    Interface list follows Appendix A (97 interfaces, no channel-channel ones);
    the exact pairing is synthetic (the coupling workbook is not parsed here).
    """
    M, nsc, nst = 13, 6, 18
    S = nsc + nst + 1
    jk = S - 1
    fs: List[Tuple[int, int]] = []
    for c in range(12):
        fs.append((c, c // 2))                        # CHAN - STR_SC (12)
        for k in range(3):
            fs.append((c, nsc + (3 * (c // 2) + k) % nst))  # CHAN - STR_STAB (36)
        fs.append((c, jk))                            # CHAN - Z_JACKET (12)
    ss: List[Tuple[int, int]] = []
    for k in range(nst):
        ss.append((nsc + k, nsc + (k + 1) % nst) if k + 1 < nst else (nsc, nsc + k))
    for k in range(nsc):
        ss.append((k, nsc + 3 * k))                   # STR_SC - STR_STAB (6)
        ss.append((k, jk))                            # STR_SC - Z_JACKET (6)
        ss.append((nsc + 3 * k + 1, jk))              # STR_STAB - Z_JACKET (6)
    t = Topology(
        n_ch=M, n_sol=S,
        ch_area=(6.5e-7,) * 12 + (1.963e-5,), ch_dh=(2.9115e-4,) * 12 + (2.499e-3,),
        ch_forward=(True,) * M, ch_intial=(3,) * M,
        sol_area=(1.28e-5,) * nsc + tuple((8.608e-6, 1.13187e-5, 2.10568e-5)[k % 3] for k in range(nst)) + (1.3972e-4,),
        sol_costeta=(1.0,) * S,
        ff=(), fs=tuple(fs), ss=tuple(ss), es=(jk,), es_choice2=(False,),
    )
    return replace(t, **kw)


# ----------------------------------------------------------------------------
# Synthetic states.
# ----------------------------------------------------------------------------

def quantize26(x: np.ndarray) -> np.ndarray:
    """Round to 26 significant bits so that x*x is exact in binary64.

    The reference squares some scalars with ``np.float64 ** 2.`` which goes
    through libm ``pow`` and is not correctly rounded for every input
    (measured here: 85 of 1e5 random doubles differ from x*x by one ulp).  On
    inputs quantised like this ``pow(x, 2.)`` is exact, which lets the golden
    fixtures pin the oracle bit-for-bit.
    """
    x = np.asarray(x, dtype=np.float64)
    m, e = np.frexp(x)
    return np.ldexp(np.round(m * (1 << 26)) / (1 << 26), e)


def synthetic_step_arrays(
    topo: Topology,
    n_nodes: int,
    batch: int | None = None,
    seed: int = 20261019,
    length: float = 10.0,
    dt: float = 0.1,
    num_step: int = 2,
    refined_mesh: bool = False,
    exact_squares: bool = True,
    heated: bool = True,
) -> StepArrays:
    """A physically plausible, randomly perturbed state for `topo`.

    Coolant-like magnitudes follow supercritical helium at 4.5 K / 6 bar
    (Appendix A of SURVEY.md); the numbers are synthetic -- CoolProp is absent.
    """
    topo.validate()
    rng = np.random.default_rng(seed)
    M, S, d = topo.n_ch, topo.n_sol, topo.ndf
    N, E = n_nodes, n_nodes - 1
    lead = () if batch is None else (batch,)

    def U(lo, hi, shape):
        return rng.uniform(lo, hi, lead + shape)

    if refined_mesh:
        w = rng.uniform(0.2, 1.8, E)
        dz1 = length * w / w.sum()
    else:
        z = np.linspace(0.0, length, N)
        dz1 = z[1:] - z[:-1]
    dz = np.broadcast_to(dz1, lead + (E,)).copy()

    # nodal state --------------------------------------------------------
    sgn = np.where(np.array(topo.ch_forward), 1.0, -1.0)
    xi = np.linspace(0.0, 1.0, N)
    v_nod = sgn[:, None] * U(0.5, 1.5, (M, 1)) * (1.0 + 0.02 * np.sin(6.0 * xi) + U(-0.01, 0.01, (M, N)))
    p_in = U(5.9e5, 6.1e5, (M, 1))
    p_nod = p_in - sgn[:, None] * 1.0e4 * (xi if True else 0) * U(0.8, 1.2, (M, 1)) + U(-20.0, 20.0, (M, N))
    t_nod = U(4.4, 4.7, (M, 1)) + 0.3 * np.exp(-((xi - 0.2) / 0.08) ** 2) * U(0.0, 1.0, (M, 1)) + U(0.0, 0.01, (M, N))
    ts_nod = U(4.4, 4.7, (S, 1)) + 0.5 * np.exp(-((xi - 0.2) / 0.08) ** 2) * U(0.0, 1.0, (S, 1)) + U(0.0, 0.01, (S, N))
    rho_nod = 140.0 - 8.0 * (t_nod - 4.5) + 2.0e-5 * (p_nod - 6.0e5)

    def gauss(a):
        return (a[..., :-1] + a[..., 1:]) / 2.0

    fl_gauss = np.empty(lead + (N_FL_GAUSS, M, E))
    vg, pg, tg = gauss(v_nod), gauss(p_nod), gauss(t_nod)
    rho_g = 140.0 - 8.0 * (tg - 4.5) + 2.0e-5 * (pg - 6.0e5)
    c0 = 220.0 + 15.0 * (tg - 4.5) + U(-1.0, 1.0, (M, E))
    if exact_squares:
        vg = quantize26(vg)
        c0 = quantize26(c0)
    fl_gauss[..., FL_V, :, :] = vg
    fl_gauss[..., FL_P, :, :] = pg
    fl_gauss[..., FL_T, :, :] = tg
    fl_gauss[..., FL_RHO, :, :] = rho_g
    fl_gauss[..., FL_C0, :, :] = c0
    fl_gauss[..., FL_GRUN, :, :] = 0.9 + 0.1 * (tg - 4.5) + U(-0.01, 0.01, (M, E))
    fl_gauss[..., FL_H, :, :] = 1.2e4 + 4.0e3 * (tg - 4.5) + U(-10.0, 10.0, (M, E))
    fl_gauss[..., FL_CV, :, :] = 2.5e3 + 200.0 * (tg - 4.5) + U(-5.0, 5.0, (M, E))
    fl_gauss[..., FL_F, :, :] = 0.02 * U(0.9, 1.1, (M, E))

    fl_node = np.empty(lead + (N_ND, M))
    fl_node[..., ND_VFIRST, :] = v_nod[..., 0]
    fl_node[..., ND_VLAST, :] = v_nod[..., -1]
    fl_node[..., ND_RHOFIRST, :] = rho_nod[..., 0]
    fl_node[..., ND_RHOLAST, :] = rho_nod[..., -1]

    fl_bc = np.empty(lead + (N_BC, M))
    area = np.array(topo.ch_area)
    fl_bc[..., BC_PINL, :] = np.where(sgn > 0, p_nod[..., 0], p_nod[..., -1])
    fl_bc[..., BC_POUT, :] = np.where(sgn > 0, p_nod[..., -1], p_nod[..., 0])
    fl_bc[..., BC_TINL, :] = U(4.3, 4.7, (M,))
    fl_bc[..., BC_TOUT, :] = U(4.3, 4.7, (M,))
    mdt = np.abs(area * v_nod[..., 0] * rho_nod[..., 0])
    fl_bc[..., BC_MDTIN, :] = mdt * sgn
    fl_bc[..., BC_MDTOUT, :] = mdt * sgn * U(0.98, 1.02, (M,))

    sol_gauss = np.empty(lead + (N_SOL_GAUSS, S, E))
    tsg = gauss(ts_nod)
    sol_gauss[..., SOL_RHO, :, :] = U(7000.0, 9000.0, (S, 1)) * np.ones(E)
    sol_gauss[..., SOL_CP, :, :] = 0.1 * tsg ** 3 * U(0.8, 1.2, (S, 1)) + 0.05
    sol_gauss[..., SOL_K, :, :] = U(50.0, 400.0, (S, 1)) * (tsg / 4.5)

    sol_q = np.zeros(lead + (2, S, E, 2))
    if heated:
        zc = np.cumsum(dz1) - dz1 / 2.0
        win = ((zc > 0.1 * length) & (zc < 0.3 * length)).astype(np.float64)
        q0 = U(50.0, 1000.0, (1, 1))
        for l in range(0, S, max(1, S // 2)):
            sol_q[..., 0, l, :, 0] = q0[..., 0, :] * win
            sol_q[..., 1, l, :, 0] = q0[..., 0, :] * win
            sol_q[..., 0, l, :, 1] = 0.5 * q0[..., 0, :] * win
            sol_q[..., 1, l, :, 1] = 0.5 * q0[..., 0, :] * win

    nff, nfs, nss, nes = len(topo.ff), len(topo.fs), len(topo.ss), len(topo.es)
    peri_ff = np.empty(lead + (2, nff, E))
    peri_ff[..., 0, :, :] = U(0.005, 0.01, (nff, 1)) * np.ones(E)
    peri_ff[..., 1, :, :] = U(0.01, 0.03, (nff, 1)) * np.ones(E)
    htc_ff = U(300.0, 1500.0, (2, nff, E))
    peri_fs = U(0.01, 4.0, (nfs, 1)) * np.ones(E)
    htc_fs = U(200.0, 2000.0, (nfs, E))
    peri_ss = U(0.005, 0.05, (nss, 1)) * np.ones(E)
    htc_ss = U(100.0, 1000.0, (nss, E))
    peri_es = U(0.05, 0.2, (nes, 1)) * np.ones(E)
    htc_es = U(1.0, 10.0, (nes, 3, E))
    qsource = np.zeros(lead + (N, S))

    sysvar = np.empty(lead + (N, d))
    sysvar[..., :, 0:M] = np.swapaxes(v_nod, -1, -2)
    sysvar[..., :, M:2 * M] = np.swapaxes(p_nod, -1, -2)
    sysvar[..., :, 2 * M:3 * M] = np.swapaxes(t_nod, -1, -2)
    sysvar[..., :, 3 * M:] = np.swapaxes(ts_nod, -1, -2)
    sysvar = sysvar.reshape(lead + (N * d,))
    syslod = np.zeros(lead + (N * d, 2))
    if num_step > 1:
        syslod[..., 0] = U(-1.0, 1.0, (N * d,)) * 1.0e-2

    return StepArrays(
        dz=dz, fl_gauss=fl_gauss, fl_node=fl_node, fl_bc=fl_bc, sol_gauss=sol_gauss,
        sol_q=sol_q, peri_ff=peri_ff, htc_ff=htc_ff, peri_fs=peri_fs, htc_fs=htc_fs,
        peri_ss=peri_ss, htc_ss=htc_ss, peri_es=peri_es, htc_es=htc_es, qsource=qsource,
        sysvar=sysvar, syslod=syslod, t_env=300.0, dt=dt, num_step=num_step,
    )


def take(arr: StepArrays, b: int) -> StepArrays:
    """Slot `b` of a batched StepArrays as an unbatched one."""
    kw = {}
    for f in fields(arr):
        v = getattr(arr, f.name)
        kw[f.name] = v[b].copy() if isinstance(v, np.ndarray) else v
    return StepArrays(**kw)
