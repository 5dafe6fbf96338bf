"""Shared set-up of the transient tests: decks from their committed image (tests/golden/decks_summary.json -- the GPU
box has no workbooks), the coolant tables of each case, the oracle's time loop."""
import json
import os

import numpy as np

from helpers import GOLDEN_DIR

CASES = {"case1": "CASE_1_ITER_like_LTS", "case1_current": "CASE_1_ITER_like_LTS", "case2": "CASE_2_ENEA_HTS_CICC",
         "case3": "CASE_3_HTS_HVDC"}
_SUMMARY = None


def deck_of(tag):
    from opensc2_b200.deck import deck_from_dict

    global _SUMMARY
    if _SUMMARY is None:
        _SUMMARY = json.load(open(os.path.join(GOLDEN_DIR, "decks_summary.json")))
    return deck_from_dict(_SUMMARY[CASES[tag]])


def tables_of(tag):
    """Coolant tables per fluid index: helium is a stand-in everywhere (CoolProp absent), nitrogen is
    the validated Span EOS table."""
    from opensc2_b200.model import helium_gas_table, synthetic_helium_table

    if tag == "case3":
        from opensc2_b200.eos import build_table
        from opensc2_b200.eos import nitrogen as N2
        return [helium_gas_table(), build_table(N2.state, (2.0e5, 4.0e5), (90.0, 160.0), 201, 281)]
    return [synthetic_helium_table()]


def sweep_kwargs(tag):
    if tag == "case1_current":
        return dict(q0=np.array([600.0]), b_field=np.array([[5.5, 5.8]]), i_op=np.array([30.0e3]), eps=np.array([-0.0035]))
    return {}


def legacy_of(tag):
    return tag in ("case2", "case3")


def field_drift(a, b, n_nodes, topo):
    """max over the nodes of |a - b| per equation, normalised by max |b| of that equation: (d,)."""
    d = topo.ndf
    a, b = a.reshape(n_nodes, d), b.reshape(n_nodes, d)
    return np.max(np.abs(a - b), axis=0) / np.maximum(np.max(np.abs(b), axis=0), 1e-300)


def oracle_transient(tr, n_steps, tcs=None, snapshot_every=0):
    """The CPU oracle's time loop on conductor 0..B-1 of a `runner.Transient` set-up (same initial state, tables,
    sweep, geometry): operating conditions + step() per time step, like the device.  Returns {step: sysvar}."""
    import dataclasses

    from oracle import oracle as O

    topo, model = tr.deck.topology, tr.deck.model
    N = tr.zcoord.shape[0]
    arr = tr.geometry.copy()
    arr.sysvar, arr.syslod = tr.sysvar0.copy(), np.zeros((tr.batch, N * topo.ndf, 2))
    sweep = dataclasses.replace(tr.sweep, tcs=tcs)
    dt = float(tr.deck.transient["STPMIN"])
    arr.dt = dt
    snaps, t = {}, 0.0
    for k in range(1, n_steps + 1):
        t = t + dt
        ins = O.operating_conditions(topo, model, tr.tables, sweep, arr.sysvar, N, t, n_threads=4, peri_ss_nodal=arr.peri_ss_nodal)
        for key, v in ins.items():
            setattr(arr, key, v)
        arr.num_step = k
        out = O.step(topo, arr, n_threads=4)
        assert np.all(out["status"] == 0), f"oracle: singular system at step {k}"
        arr.sysvar, arr.syslod = out["sysvar"], out["syslod"]
        if (snapshot_every and k % snapshot_every == 0) or k == n_steps:
            snaps[k] = arr.sysvar.copy()
    return snaps


def helium_table_from_golden(tag, chans=None, p_pad=2.0e3, t_pad=0.02, n_p=401, n_t=41, t_props=None):
    """A LOCAL helium table rebuilt from the property columns of the reference's own shipped result (CoolProp outputs along
    the end state of the case), for the comparison of a steady state with that result -- CoolProp and its helium equation
    of state are not available here.  The samples of all channels lie on a line T_l(p) of the (p, T) plane;
    every property is interpolated along that line in p and continued off it to first order in (T - T_l(p)) where
    thermodynamics gives the slope: d rho / dT = -rho beta, dh / dT = cp, with beta = Gruneisen cp / w^2 and
    kappa = cp / (cv rho w^2) from the columns; the other properties are taken constant across the few mK the states
    visited here are away from the line.  Valid only within `t_pad` of the line: a transient with a heat slug leaves it."""
    from opensc2_b200.model import (FP_BETA, FP_CP, FP_CV, FP_H, FP_K, FP_KAPPA, FP_MU, FP_PRANDTL, FP_RHO, FP_SOUND,
                                    FluidTable)

    z = np.load(os.path.join(GOLDEN_DIR, "tdd_solutions.npz"))
    chans = chans or [k.split("/")[1] for k in z.files if k.startswith(f"{tag}/CHAN") and k.endswith("/all")]
    rows = np.concatenate([z[f"{tag}/{c}/all"] for c in chans])
    rows = rows[np.argsort(rows[:, 1])]
    _, first = np.unique(rows[:, 1], return_index=True)
    rows = rows[first]
    if rows.shape[1] == 15:       # legacy files (CASE_1, CASE_2): x p T rho v mu Gruneisen h cp cv w k Re Pr mdot
        p, T, rho, mu, phi, h, cp, cv, w, k, pr = (rows[:, i] for i in (1, 2, 3, 5, 6, 7, 8, 9, 10, 11, 13))
        kappa, beta = cp / (cv * rho * w ** 2), phi * cp / w ** 2
    else:                         # HEAD files (CASE_3): x p T rho v beta kappa Pr mu h cp cv w k Re Gruneisen mdot
        p, T, rho, beta, kappa, pr, mu, h, cp, cv, w, k = (rows[:, i] for i in (1, 2, 3, 5, 6, 7, 8, 9, 10, 11, 12, 13))
    pg = np.linspace(p.min() - p_pad, p.max() + p_pad, n_p)
    tg = np.linspace(T.min() - t_pad, T.max() + t_pad, n_t)
    on = lambda col: np.interp(pg, p, col)                        # along the line (held constant beyond its ends)
    dT = tg[None, :] - on(T)[:, None]
    # the density column belongs to the end state (T above); the other property columns were written from the nodal
    # pass of the step before -- at a steady state that is the same temperature, except where the caller knows
    # better (`t_props`: CASE_3's helium columns carry the uniform initial 30 K, not the 30 - 31.2 K of its end state)
    dTp = dT if t_props is None else (tg[None, :] - t_props) * np.ones((n_p, 1))
    props = np.empty((10, n_p, n_t))
    flat = lambda col: np.repeat(on(col)[:, None], n_t, axis=1)
    props[FP_RHO] = on(rho)[:, None] * (1.0 - on(beta)[:, None] * dT)
    props[FP_H] = on(h)[:, None] + on(cp)[:, None] * dTp
    for idx, col in ((FP_BETA, beta), (FP_KAPPA, kappa), (FP_PRANDTL, pr), (FP_MU, mu), (FP_CP, cp), (FP_CV, cv), (FP_SOUND, w), (FP_K, k)):
        props[idx] = flat(col)
    return FluidTable(p_min=float(pg[0]), p_step=float(pg[1] - pg[0]), t_min=float(tg[0]), t_step=float(tg[1] - tg[0]), props=props)
