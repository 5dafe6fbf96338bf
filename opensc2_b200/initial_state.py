"""Initial state of a conductor: what the reference computes once, in
`Conductor.initialization` (/root/reference/source_code/conductor.py:2679-2944), before the first time step.

Host-side (numpy), once per run; nothing here is on the time-step path.
  * `gen_flow`           -- utility_functions/gen_flow.py:8-47, 72-262, 263-395, 399-836: the initial flow of every
                            channel from its INTIAL mode (1: p_in, p_out -> mdot; 2: p_in, mdot_out -> p_out;
                            3: p_out, mdot_in -> p_in), stand-alone channels and groups in hydraulic parallel;
                            it OVERWRITES the boundary values (PREINL / PREOUT / MDTIN / MDTOUT) the step imposes.
  * `coolant_initial_state` -- coolant.py:87-130: linear p and T between inlet and outlet, rho(p, T), v = mdot / (A rho).
  * `solid_initial_temperature` -- utility_functions/solid_components_initialization.py:6-79.
  * `pack_sysvar`        -- conductor.py:2911-2944.
Coolant properties: `PropsSI` of the reference = bilinear lookup in the same (p, T) tables the device uses.
Only positive INTIAL flags: the negative ones read time-dependent boundary values from an external workbook
(`get_from_xlsx`), which is refused, not silently replaced.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

from .correlations import channel_friction_factor
from .model import FP_MU, FP_RHO, FluidTable, Model, table_lookup
from .problem import BC_MDTIN, BC_MDTOUT, BC_PINL, BC_POUT, BC_TINL, BC_TOUT, Topology

MAX_ITER = 1000      # gen_flow.py:17
TOL = 1.0e-10        # gen_flow.py:18


def _rho_mu(tab: FluidTable, p: float, t: float):
    """Coolant.eval_coolant_density_din_viscosity_gen_flow (coolant.py:54-67)."""
    return (float(table_lookup(tab, FP_RHO, np.array([p]), np.array([t]))[0]),
            float(table_lookup(tab, FP_MU, np.array([p]), np.array([t]))[0]))


def _friction(model, topo, j, re) -> float:
    return float(channel_friction_factor(model, topo, j, np.array([re]), nodal=None)[0])      # gen_flow: nodal = None


def hydraulic_groups(topo: Topology, peri_ff_open: np.ndarray) -> (List[List[int]], List[int]):
    """Channels joined by an OPEN contact perimeter share inlet and outlet pressure (hydraulic parallel,
    conductor.py:1354-1494, 1908-2103); everything else is initialised on its own."""
    parent = list(range(topo.n_ch))

    def find(a):
        while parent[a] != a:
            a = parent[a]
        return a

    for k, (a, b) in enumerate(topo.ff):
        if peri_ff_open[k] > 0.0:
            ra, rb = find(a), find(b)
            if ra != rb:
                parent[max(ra, rb)] = min(ra, rb)
    groups = {}
    for j in range(topo.n_ch):
        groups.setdefault(find(j), []).append(j)
    parallel = [g for g in groups.values() if len(g) > 1]
    alone = [g[0] for g in groups.values() if len(g) == 1]
    return parallel, alone


def _velocity_from_pressure_drop(model, topo, j, length, costeta, pdrop, rho, mu) -> float:
    """Coolant.compute_velocity_gen_flow (coolant.py:265-319)."""
    leff = length / costeta
    err, f, v, ii = 1.0, 0.01, 0.0, 1
    dh = topo.ch_dh[j]
    while ii <= MAX_ITER and err >= TOL:
        ii += 1
        old = v
        v = np.sqrt((dh * pdrop) / (2.0 * f * rho * leff))
        re = (rho * v * dh) / mu
        f = _friction(model, topo, j, re)
        err = np.fabs(v - old) / np.fabs(v)
    return float(v)


def gen_flow(topo: Topology, model: Model, tables: Sequence[FluidTable], fl_bc: np.ndarray, length: float,
             peri_ff_open: np.ndarray, ch_costeta: Sequence[float] | None = None) -> np.ndarray:
    """Returns the (6, M) boundary values after the reference's overwrites.  `fl_bc` rows: PREINL, PREOUT,
    TEMINL, TEMOUT, MDTIN, MDTOUT as read from the deck (mass flow rates positive)."""
    bc = np.array(fl_bc, dtype=np.float64, copy=True)
    M = topo.n_ch
    cos = [1.0] * M if ch_costeta is None else list(ch_costeta)
    sgn = [1.0 if f else -1.0 for f in topo.ch_forward]          # channel.flow_dir[1]
    parallel, alone = hydraulic_groups(topo, peri_ff_open)

    def g0(j):
        return 2.0 * length / (topo.ch_dh[j] * (topo.ch_area[j] ** 2))

    def reynolds(j, mdot, mu):                                     # coolant.py:69-85
        return mdot * topo.ch_dh[j] / (mu * topo.ch_area[j])

    # ---- stand-alone channels: initialize_flow_no_hydraulic_parallel (gen_flow.py:72-262) -----------------
    for j in alone:
        tab = tables[model.ch_fluid[j]]
        mode = topo.ch_intial[j]
        if mode == 1:
            p_inl, p_out, t_inl = bc[BC_PINL, j], bc[BC_POUT, j], bc[BC_TINL, j]
            rho, mu = _rho_mu(tab, (p_inl + p_out) / 2, t_inl)
            delta_p = p_inl - p_out
            if delta_p < 0:
                raise ValueError(f"Error in function gen_flow!\nNegative pressure drop: delta_p = {delta_p}")
            v = _velocity_from_pressure_drop(model, topo, j, length, cos[j], delta_p, rho, mu)
            mdot = sgn[j] * topo.ch_area[j] * rho * v              # compute_mass_flow_with_direction
            if abs(mdot) != bc[BC_MDTIN, j]:
                bc[BC_MDTIN, j] = mdot
        else:
            # get_missing_pressure_no_hydraulic_parallel (gen_flow.py:263-395)
            mdot_known = bc[BC_MDTOUT, j] if mode == 2 else bc[BC_MDTIN, j]
            p_known = bc[BC_PINL, j] if mode == 2 else bc[BC_POUT, j]
            t_inl, t_out = bc[BC_TINL, j], bc[BC_TOUT, j]
            rho_k, mu_k = _rho_mu(tab, p_known, t_inl)
            f = _friction(model, topo, j, reynolds(j, mdot_known, mu_k))
            dp_old = float(g0(j) * f * mdot_known ** 2 / rho_k)
            t_ave = (t_inl + t_out) / 2
            err, it, p_missing = 10.0, 0, p_known
            while err >= TOL and it < MAX_ITER:
                it += 1
                p_missing = p_known - dp_old if mode == 2 else p_known + dp_old
                p_ave = (p_known + p_missing) / 2.0
                rho_a, mu_a = _rho_mu(tab, p_ave, t_ave)
                f = _friction(model, topo, j, reynolds(j, mdot_known, mu_a))
                dp_new = float(g0(j) * f * mdot_known ** 2 / rho_a)
                err = abs(dp_old - dp_new) / dp_old
                dp_old = dp_new
            bc[BC_POUT if mode == 2 else BC_PINL, j] = p_missing
            bc[BC_MDTIN, j] = sgn[j] * bc[BC_MDTIN, j]             # gen_flow.py:389-391

    # ---- groups in hydraulic parallel (gen_flow.py:399-836) -----------------------------------------------
    for grp in parallel:
        modes = {topo.ch_intial[j] for j in grp}
        if len(modes) != 1:
            raise ValueError("ERROR: channels in hydraulic parallel must share abs(INTIAL) (gen_flow.check_intial_values)")
        dirs = {topo.ch_forward[j] for j in grp}
        if len(dirs) != 1:
            raise ValueError("Channels in hydraulic parallel must all have the same flow direction.")
        mode, n = modes.pop(), len(grp)
        if mode == 1:                                              # abs_intial_equal_1_hp (:512-642)
            p_inl_ave = np.sum(bc[BC_PINL, grp]) / n
            p_out_ave = np.sum(bc[BC_POUT, grp]) / n
            p_ave = (p_inl_ave + p_out_ave) / 2.0
            delta_p = p_inl_ave - p_out_ave
            if delta_p < 0:
                raise ValueError(f"Error in function gen_flow!\nNegative pressure drop: delta_p = {delta_p}")
            for j in grp:
                tab = tables[model.ch_fluid[j]]
                rho, mu = _rho_mu(tab, p_ave, bc[BC_TINL, j])
                v = _velocity_from_pressure_drop(model, topo, j, length, cos[j], delta_p, rho, mu)
                mdot = sgn[j] * topo.ch_area[j] * rho * v
                if abs(mdot) != bc[BC_MDTIN, j]:
                    bc[BC_MDTIN, j] = mdot
                bc[BC_PINL, j], bc[BC_POUT, j] = p_inl_ave, p_out_ave
        else:                                                      # abs_intial_equal_2_or_3_hp (:648-829)
            key_a, key_b, key_m = (BC_PINL, BC_POUT, BC_MDTOUT) if mode == 2 else (BC_POUT, BC_PINL, BC_MDTIN)
            mdot_known = bc[key_m, grp].copy()
            p_ave = np.sum(bc[key_a, grp]) / n
            g = np.array([g0(j) for j in grp])
            rho, fric = np.zeros(n), np.zeros(n)
            for i, j in enumerate(grp):
                rho[i], mu = _rho_mu(tables[model.ch_fluid[j]], p_ave, bc[BC_TINL, j])
                fric[i] = _friction(model, topo, j, reynolds(j, mdot_known[i], mu))
            alpha = g * fric / rho
            mdot_group = np.sum(mdot_known)
            dp_group = (abs(mdot_group) / np.sum(1 / np.sqrt(alpha))) ** 2
            mdot_new = np.sqrt(dp_group * rho / (g * fric))
            p_missing = p_ave - dp_group if mode == 2 else p_ave + dp_group
            for i, j in enumerate(grp):
                bc[key_b, j] = p_missing
                if mdot_new[i] != bc[key_m, j]:
                    bc[key_m, j] = sgn[j] * mdot_new[i]
    return bc


def coolant_initial_state(topo: Topology, model: Model, tables: Sequence[FluidTable], bc: np.ndarray,
                          zcoord: np.ndarray, length: float):
    """Nodal (v, p, T, rho) of every channel: Coolant._eval_nodal_pressure_temperature_velocity_initialization
    (coolant.py:87-130).  Returns arrays of shape (M, N)."""
    M, N = topo.n_ch, zcoord.shape[0]
    v, p, t, rho = (np.zeros((M, N)) for _ in range(4))
    for j in range(M):
        mode = topo.ch_intial[j]
        mfr = bc[BC_MDTOUT, j] if mode == 2 else bc[BC_MDTIN, j]
        if mfr >= 0.0:
            p[j] = np.interp(zcoord, [0.0, length], [bc[BC_PINL, j], bc[BC_POUT, j]])
        else:
            p[j] = np.interp(zcoord, [0.0, length], [bc[BC_POUT, j], bc[BC_PINL, j]])
        t[j] = np.interp(zcoord, [0.0, length], [bc[BC_TINL, j], bc[BC_TOUT, j]])
        rho[j] = table_lookup(tables[model.ch_fluid[j]], FP_RHO, p[j], t[j])
        v[j] = mfr / (np.maximum(topo.ch_area[j], 1e-7) * rho[j])
    return v, p, t, rho


def solid_initial_temperature(topo: Topology, solid_ops, peri_fs: np.ndarray, t_chan: np.ndarray,
                              zcoord: np.ndarray, length: float) -> np.ndarray:
    """(S, N) solid temperatures: INTIAL 0 -> contact-perimeter weighted channel temperature (minimum channel
    temperature when the solid touches no channel), INTIAL 1 -> linear TEMINL..TEMOUT
    (solid_components_initialization.py:6-79)."""
    S, N, M = topo.n_sol, zcoord.shape[0], topo.n_ch
    out = np.zeros((S, N))
    t_min = t_chan.min(axis=1)
    for l in range(S):
        mode = int(solid_ops[l].get("INTIAL", 0))
        if mode == 0:
            weight = np.zeros(M)
            for k, (j, ll) in enumerate(topo.fs):
                if ll == l:
                    weight[j] = peri_fs[k]
            if np.sum(weight) > 0:
                for j in range(M):
                    out[l] = out[l] + t_chan[j] * weight[j] / np.sum(weight)
            else:
                out[l] = np.ones(N) * t_min.min()
        elif mode == 1:
            out[l] = np.interp(zcoord, [0.0, length], [solid_ops[l]["TEMINL"], solid_ops[l]["TEMOUT"]])
        else:
            raise ValueError(f"solid {l}: INTIAL {mode} is not available (the reference: 'still to do')")
    return out


def pack_sysvar(v: np.ndarray, p: np.ndarray, t: np.ndarray, ts: np.ndarray) -> np.ndarray:
    """SYSVAR[:, 0] (conductor.py:2911-2944): node-major, equations v_1..v_M, p_1..p_M, T_1..T_M, Ts_1..Ts_S."""
    M, N = v.shape
    S = ts.shape[0]
    d = 3 * M + S
    x = np.empty((N, d))
    x[:, 0:M], x[:, M:2 * M], x[:, 2 * M:3 * M], x[:, 3 * M:] = v.T, p.T, t.T, ts.T
    return x.reshape(N * d)


@dataclass
class InitialState:
    zcoord: np.ndarray       # (N,)
    fl_bc: np.ndarray        # (6, M) after gen_flow
    sysvar: np.ndarray       # (N * d,)
    rho: np.ndarray          # (M, N) nodal coolant density at t = 0


def initial_state(deck, tables: Sequence[FluidTable], zcoord: np.ndarray, fl_bc: np.ndarray | None = None,
                  legacy_intial2: bool = False) -> InitialState:
    """Deck -> initial state of ONE conductor; `fl_bc` overrides the deck's boundary values (sweeps).

    `legacy_intial2`: the shipped TDD_examples decks say INTIAL = 2 in an older meaning -- inlet pressure, inlet
    temperature and mass flow are known, gen_flow computes the OUTLET pressure, and the transient then imposes inlet
    mass flow + that outlet pressure (HEAD's INTIAL = 3 rows).  Evidence: golden CASE_3 CHAN_2 keeps p(outlet) =
    260 115.741 Pa = what gen_flow's INTIAL = 2 branch yields from PREINL = 3.0e5 Pa with CoolProp-exact nitrogen
    properties (profiles/r2_coolant_table_validation.txt), while PREOUT of the deck is 2.6e5.  With the flag, the flow
    is generated with HEAD's INTIAL = 2 branch (MDTOUT := MDTIN) and everything after it uses INTIAL = 3."""
    from dataclasses import replace

    topo, model = deck.topology, deck.model
    bc0 = deck.fl_bc if fl_bc is None else fl_bc
    peri_open = deck.peri["ff"][0] if len(topo.ff) else np.zeros(0)
    if legacy_intial2:
        if any(i != 3 for i in topo.ch_intial):
            raise ValueError("legacy_intial2 expects the deck loaded with intial_override=3")
        bc0 = np.array(bc0, dtype=np.float64, copy=True)
        bc0[BC_MDTOUT] = bc0[BC_MDTIN]
        bc = gen_flow(replace(topo, ch_intial=(2,) * topo.n_ch), model, tables, bc0, deck.length, peri_open)
        bc[BC_MDTOUT] = bc[BC_MDTIN]                # signed like MDTIN from here on
    else:
        bc = gen_flow(topo, model, tables, bc0, deck.length, peri_open)
    v, p, t, rho = coolant_initial_state(topo, model, tables, bc, zcoord, deck.length)
    ts = solid_initial_temperature(topo, deck.solid_ops, deck.peri["fs"], t, zcoord, deck.length)
    return InitialState(zcoord=zcoord, fl_bc=bc, sysvar=pack_sysvar(v, p, t, ts), rho=rho)
