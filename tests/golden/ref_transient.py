"""Full transients run by the UNMODIFIED reference functions in a loop (build container only), the pin of
`opensc2_b200.runner.Transient` over 1000-2000 chained steps (VERDICT r1 "next" 1a).

Per time step, exactly the sequence of Simulation.conductor_solution (simulation.py:341-445) with no current defined:
    operating_conditions_em  -> Conductor.__eval_gauss_point_em            |  make_golden_opcond.
    operating_conditions_th  -> Conductor.eval_transp_coeff (Gauss pass)   |  reference_operating_conditions
    build_heat_source        -> SolidComponent.get_heat + Gauss-point part |  (reference objects, reference code)
    step(conductor, environment, qsource, num_step)                        |  ref_harness.run_reference_step
`PropsSI` := bilinear lookup in the same tables the device reads; Tcs from the reference's own
current_sharing_temperature_nb3sn (scipy bisect) once at step 0.  Initial state and mesh: opensc2_b200.initial_state /
mesh (restatements; their pins are the golden mesh coordinates and the golden outlet pressure of CASE_3 CHAN_2).

Writes tests/golden/transient_<tag>.npz: snapshots of SYSVAR every `every` steps + everything needed to rebuild the
run on the device.  Usage:  python tests/golden/ref_transient.py case1 | case1_current | case2 | case3
"""
import os
import sys
import time as _time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import ref_harness as H  # noqa: E402
from make_golden_opcond import reference_operating_conditions  # noqa: E402

from opensc2_b200.deck import load_deck  # noqa: E402
from opensc2_b200.initial_state import initial_state  # noqa: E402
from opensc2_b200.mesh import build_mesh  # noqa: E402
from opensc2_b200.runner import deck_sweep, step_geometry  # noqa: E402
from opensc2_b200.problem import take  # noqa: E402

DECKS = "/root/reference/TDD_examples"


def case_setup(tag):
    """(deck, tables, fluid_names, kwargs of the sweep, legacy flag, n_steps, snapshot interval)."""
    from opensc2_b200.eos import build_table
    from opensc2_b200.eos import nitrogen as N2
    from opensc2_b200.model import helium_gas_table, synthetic_helium_table

    if tag in ("case1", "case1_current"):
        deck = load_deck(os.path.join(DECKS, "CASE_1_ITER_like_LTS"), intial_override=1)
        sw = {} if tag == "case1" else dict(q0=np.array([600.0]), b_field=np.array([[5.5, 5.8]]), i_op=np.array([30.0e3]),
                                             eps=np.array([-0.0035]))
        return deck, [synthetic_helium_table()], ["helium"], sw, False, (1000 if tag == "case1" else 300), 100
    if tag == "case3":
        deck = load_deck(os.path.join(DECKS, "CASE_3_HTS_HVDC"), intial_override=3)
        tabs = [helium_gas_table(), build_table(N2.state, (2.0e5, 4.0e5), (90.0, 160.0), 201, 281)]
        return deck, tabs, ["helium", "nitrogen"], {}, True, 2000, 100
    if tag == "case2":
        deck = load_deck(os.path.join(DECKS, "CASE_2_ENEA_HTS_CICC"), intial_override=3)
        return deck, [synthetic_helium_table()], ["helium"], {}, True, int(os.environ.get("OSC2_REF_CASE2_STEPS", 320)), 80
    raise SystemExit(f"unknown case {tag}")


def reference_tcs(deck, sweep, zcoord):
    """(1, S, E) T_cur_sharing_min at the Gauss points by the reference's own function."""
    H._import_reference()
    from properties_of_materials.niobium3_tin import current_sharing_temperature_nb3sn

    S, N = deck.topology.n_sol, zcoord.shape[0]
    tcs = np.zeros((1, S, N - 1))
    for l, sm in enumerate(deck.model.solids):
        if sm.kind != 0:
            continue
        bn = np.linspace(sweep.b_field[0, l, 0], sweep.b_field[0, l, 1], N)
        bg = np.abs(bn[:-1] + bn[1:]) / 2.0
        eps = np.full(N - 1, (sweep.eps[0, l] + sweep.eps[0, l]) / 2.0)
        jop = np.full(N - 1, sweep.jop[0, l])
        tcs[0, l] = current_sharing_temperature_nb3sn(bg, eps, jop, sm.tc0m, sm.bc20m, sm.c0)
    return tcs


def main(tag):
    assert H.reference_available()
    deck, tables, names, sw, legacy, n_steps, every = case_setup(tag)
    topo, model = deck.topology, deck.model
    zcoord = build_mesh(deck.length, deck.mesh)
    N = zcoord.shape[0]
    ini = initial_state(deck, tables, zcoord, legacy_intial2=legacy)
    sweep = deck_sweep(deck, zcoord, 1, **sw)
    if sweep.jop is not None:
        sweep.tcs = reference_tcs(deck, sweep, zcoord)
    geo = step_geometry(deck, zcoord, 1, ini.fl_bc[None])
    arr = take(geo, 0)
    arr.sysvar, arr.syslod = ini.sysvar.copy(), np.zeros((N * topo.ndf, 2))
    dt = float(deck.transient["STPMIN"])
    arr.dt = dt
    pn = geo.peri_ss_nodal
    snaps, t = {0: arr.sysvar.copy()}, 0.0
    t0 = _time.time()
    for k in range(1, n_steps + 1):
        t = t + dt
        ins = reference_operating_conditions(topo, model, None, sweep, arr.sysvar[None], N, t, b=0, peri_ss_nodal=pn,
                                             tables=tables, fluid_names=names, zcoord=zcoord,
                                             solid_inputs=deck.solid_inputs, ss_view_factor=deck.ss_view_factor)
        for key, v in ins.items():
            setattr(arr, key, v)
        arr.num_step = k
        out = H.run_reference_step(topo, arr)
        arr.sysvar, arr.syslod = out["sysvar"], out["syslod"]
        if k % every == 0 or k == n_steps:
            snaps[k] = arr.sysvar.copy()
            print(f"{tag}: step {k}/{n_steps}  t = {t:.3f}  ({_time.time() - t0:.0f} s)", flush=True)
    steps = np.array(sorted(snaps))
    payload = dict(steps=steps, sysvar=np.stack([snaps[k] for k in steps]), zcoord=zcoord, fl_bc=ini.fl_bc, dt=np.array(dt),
                   b_field=sweep.b_field, eps=sweep.eps, q0=sweep.q0, q_window=sweep.q_window)
    if sweep.jop is not None:
        payload.update(jop=sweep.jop, tcs=sweep.tcs)
    np.savez_compressed(os.path.join(HERE, f"transient_{tag}.npz"), **payload)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "case1")
