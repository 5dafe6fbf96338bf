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
