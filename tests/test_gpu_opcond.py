"""GPU parity of K_P (device-side operating conditions: Gauss averaging, coolant tables,
friction / Nusselt correlations, HTC models, material fits + homogenisation, heat source)
against the CPU oracle (oracle/props_oracle.c, itself pinned to the reference's fits in
tests/test_props_oracle_golden.py).

Tolerances (stated FP64 relative tolerance of north_star):
  * everything without a transcendental (Gauss means, table interpolation, Gruneisen, heat
    source, constant HTCs): bit-exact;
  * fits and correlations (chains of pow / exp / log10, CUDA libm vs glibc): RTOL_FIT = 2e-12;
  * one time step with device-side properties (K_P + step vs oracle + oracle from the same
    state): the fit differences are amplified by the conditioning of the system (cond ~ 2e8
    after row scaling): RTOL_CHAIN_V = 1e-6 on velocities, RTOL_CHAIN = 2e-7 on
    pressures and temperatures, per step."""
import numpy as np
import pytest

from helpers import assert_bits
from opensc2_b200.model import case1_model, consistent_initial_state, synthetic_helium_table, synthetic_sweep
from opensc2_b200.problem import case1_topology, synthetic_step_arrays

pytestmark = pytest.mark.gpu
RTOL_FIT = 2e-12
RTOL_CHAIN_V, RTOL_CHAIN = 1e-6, 2e-7


def rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)) if a.size else 0.0


def rel_field(a, b):
    """max |a - b| over a nodal field / max |b| of that field (per conductor and equation):
    a velocity that crosses zero has no meaningful element-wise relative error."""
    num = np.max(np.abs(a - b), axis=1)
    den = np.maximum(np.max(np.abs(b), axis=1), 1e-300)
    return float(np.max(num / den))


def _setup(correlations, n, batch, seed=5, cold_jacket=False):
    topo = case1_topology(1)
    model = case1_model(correlations)
    tab = synthetic_helium_table()
    arr = synthetic_step_arrays(topo, n, batch=batch, seed=seed)
    if cold_jacket:      # conductor 0: jacket below 12.897 K everywhere except one node above it in conductor 1
        sv = arr.sysvar.reshape(batch, n, topo.ndf)
        sv[1, n // 2, topo.ndf - 1] = 30.0
    sweep = synthetic_sweep(topo, model, n, batch, seed=seed)
    return topo, model, tab, arr, sweep


@pytest.mark.parametrize("correlations", [False, True])
@pytest.mark.parametrize("time", [5.0, 12.0])
def test_operating_conditions_match_oracle(correlations, time):
    from opensc2_b200.batch import BatchedStep
    from oracle import oracle as O

    n, batch = 65, 37
    topo, model, tab, arr, sweep = _setup(correlations, n, batch, cold_jacket=True)
    with BatchedStep(topo, batch, n) as bs:
        bs.set_model(model, [tab])
        bs.upload_sweep(sweep)
        bs.upload_state(arr.sysvar, arr.syslod)
        bs.upload_geometry(arr)
        bs.eval_operating_conditions(time)
        bs.synchronize()
        out = bs.download_step_inputs()
    ref = O.operating_conditions(topo, model, [tab], sweep, arr.sysvar, n, time)
    # exact parts
    for q in (0, 1, 2, 3, 4, 5, 6, 7):
        assert_bits(out["fl_gauss"][:, q], ref["fl_gauss"][:, q], f"fl_gauss[{q}]")
    assert_bits(out["fl_node"], ref["fl_node"], "fl_node")
    assert_bits(out["sol_q"], ref["sol_q"], "sol_q")
    assert_bits(out["sol_gauss"][:, 0], ref["sol_gauss"][:, 0], "solid density")
    assert_bits(out["htc_es"], ref["htc_es"], "htc_es")
    if not correlations:
        for key in ("htc_ff", "htc_fs", "htc_ss"):
            assert_bits(out[key], ref[key], key)
    # fits
    for key in ("fl_gauss", "sol_gauss", "htc_ff", "htc_fs", "htc_ss"):
        assert rel(out[key], ref[key]) <= RTOL_FIT, f"{key}: {rel(out[key], ref[key]):.3e}"
    # the heat slug is on at t = 12 s and off at t = 5 s
    assert (ref["sol_q"][..., 0].max() > 0.0) == (time == 12.0)
    # the whole-array SS branch differs between conductor 0 (cold jacket) and 1 (one warm node)
    assert np.max(np.abs(ref["sol_gauss"][0, 1, 1] - ref["sol_gauss"][1, 1, 1])) > 0.0


@pytest.mark.parametrize("correlations", [False, True])
def test_chained_steps_with_device_side_properties(correlations):
    """State stays on the device over [K_P -> step] x 6 (crossing the start of the heat slug).
    Per-step parity: every device step is compared with the oracle advanced ONE step from the
    device's own previous state (fit differences of 1e-15 are amplified by cond ~ 2e8 per solve,
    so a tolerance per step is the meaningful statement)."""
    from opensc2_b200.batch import BatchedStep
    from oracle import oracle as O

    n, batch, dt = 41, 9, 0.1
    topo, model, tab, _arr, sweep = _setup(correlations, n, batch, seed=11)
    arr = consistent_initial_state(topo, model, tab, n, batch, seed=11, dt=dt)
    host = arr.copy()
    d, M = topo.ndf, topo.n_ch
    worst_v = worst_pt = 0.0
    with BatchedStep(topo, batch, n) as bs:
        bs.set_model(model, [tab])
        bs.upload_sweep(sweep)
        bs.upload_state(arr.sysvar, arr.syslod)
        bs.upload_geometry(arr)
        for k in range(1, 7):
            t = 9.7 + k * dt
            bs.eval_operating_conditions(t)
            bs.step(dt, k)
            bs.synchronize()
            sv, sl, st = bs.download_state()
            assert np.all(st == 0)
            ref_in = O.operating_conditions(topo, model, [tab], sweep, host.sysvar, n, t)
            for key, v in ref_in.items():
                setattr(host, key, v)
            host.num_step, host.dt = k, dt
            ref = O.step(topo, host)
            assert np.all(ref["status"] == 0)
            a, b = sv.reshape(batch, n, d), ref["sysvar"].reshape(batch, n, d)
            worst_v = max(worst_v, rel_field(a[..., :M], b[..., :M]))
            worst_pt = max(worst_pt, rel_field(a[..., M:], b[..., M:]))
            assert rel(sl, ref["syslod"]) <= RTOL_FIT
            host.sysvar, host.syslod = sv, sl          # continue from the device's state
    print(f"per-step field-relative differences: v {worst_v:.2e}, p/T {worst_pt:.2e}")
    assert worst_v <= RTOL_CHAIN_V, worst_v
    assert worst_pt <= RTOL_CHAIN, worst_pt
    x = host.sysvar.reshape(batch, n, d)
    assert x[..., 2 * M:].max() > 4.8               # the slug heated the conductor
    assert np.all(np.isfinite(x))


def test_model_validation_rejects_unimplemented_flags():
    import dataclasses

    from opensc2_b200.batch import BatchedStep

    topo = case1_topology(1)
    bad = dataclasses.replace(case1_model(), ch_ifriction=(-99, 123))   # Petukhov: not on the device (and it raises on arrays in the reference)
    with BatchedStep(topo, 2, 9) as bs:
        with pytest.raises(ValueError):
            bs.set_model(bad, [synthetic_helium_table()])


def test_case2_flavoured_model_colebrook_aluminium_generic_kernels():
    """d = 9 (generic shared-memory kernels): IFRICTION 122 with its array-wide iteration count,
    124, Nusselt 1 / 212, aluminium + copper stabilisers, SS jacket, Kapitza-limited HTCs.
    K_P vs oracle, then three chained steps with per-step tolerance."""
    from opensc2_b200.batch import BatchedStep
    from opensc2_b200.model import case2_mini_model
    from opensc2_b200.problem import case2_mini_topology
    from oracle import oracle as O

    n, batch, dt = 33, 6, 0.05
    topo, model, tab = case2_mini_topology(), case2_mini_model(), synthetic_helium_table()
    arr = consistent_initial_state(topo, model, tab, n, batch, seed=4, dt=dt, length=3.0)
    sweep = synthetic_sweep(topo, model, n, batch, seed=4, length=3.0, with_tcs=False)
    host = arr.copy()
    d, M = topo.ndf, topo.n_ch
    with BatchedStep(topo, batch, n) as bs:
        bs.set_model(model, [tab])
        bs.upload_sweep(sweep)
        bs.upload_state(arr.sysvar, arr.syslod)
        bs.upload_geometry(arr)
        bs.eval_operating_conditions(12.0)
        bs.synchronize()
        out = bs.download_step_inputs()
        ref = O.operating_conditions(topo, model, [tab], sweep, arr.sysvar, n, 12.0)
        for key in ("fl_gauss", "sol_gauss", "htc_fs", "htc_ss", "htc_es", "fl_node", "sol_q"):
            assert rel(out[key], ref[key]) <= RTOL_FIT, f"{key}: {rel(out[key], ref[key]):.3e}"
        assert np.ptp(ref["fl_gauss"][:, 8, 0]) > 0          # Colebrook friction really varies
        worst_v = worst_pt = 0.0
        for k in range(1, 4):
            t = 11.9 + k * dt
            bs.advance(t, dt, k)
            bs.synchronize()
            sv, sl, st = bs.download_state()
            assert np.all(st == 0)
            ref_in = O.operating_conditions(topo, model, [tab], sweep, host.sysvar, n, t)
            for key, v in ref_in.items():
                setattr(host, key, v)
            host.num_step, host.dt = k, dt
            r = O.step(topo, host)
            a, b = sv.reshape(batch, n, d), r["sysvar"].reshape(batch, n, d)
            worst_v = max(worst_v, rel_field(a[..., :M], b[..., :M]))
            worst_pt = max(worst_pt, rel_field(a[..., M:], b[..., M:]))
            host.sysvar, host.syslod = sv, sl
    assert worst_v <= RTOL_CHAIN_V and worst_pt <= RTOL_CHAIN, (worst_v, worst_pt)


def test_case2_topology_full_step_on_generic_kernels():
    """CASE_2-sized system (d = 64: 13 channels, 6 RE123 + 18 Al solids, SS jacket, 97 interfaces; one
    conductor per 64-lane CTA): device-side operating conditions + step vs the oracle."""
    from opensc2_b200.batch import BatchedStep
    from opensc2_b200.model import case2_model
    from opensc2_b200.problem import case2_topology
    from oracle import oracle as O

    n, batch, dt = 9, 3, 0.05
    topo = case2_topology()
    model, tab = case2_model(topo), synthetic_helium_table()
    arr = consistent_initial_state(topo, model, tab, n, batch, seed=8, dt=dt, length=3.0)
    sweep = synthetic_sweep(topo, model, n, batch, seed=8, length=3.0, with_tcs=False)
    host = arr.copy()
    d, M = topo.ndf, topo.n_ch
    with BatchedStep(topo, batch, n) as bs:
        bs.set_model(model, [tab])
        bs.upload_sweep(sweep)
        bs.upload_state(arr.sysvar, arr.syslod)
        bs.upload_geometry(arr)
        bs.set_capture(True)
        for k in range(1, 3):
            t = 11.95 + k * dt
            bs.advance(t, dt, k)
            bs.synchronize()
            ins = bs.download_step_inputs()
            sv, sl, st = bs.download_state()
            assert np.all(st == 0)
            ref_in = O.operating_conditions(topo, model, [tab], sweep, host.sysvar, n, t)
            for key in ins:
                assert rel(ins[key], ref_in[key]) <= RTOL_FIT, f"{key}: {rel(ins[key], ref_in[key]):.3e}"
            for key, v in ref_in.items():
                setattr(host, key, v)
            host.num_step, host.dt = k, dt
            r = O.step(topo, host)
            assert np.all(r["status"] == 0)
            a, b = sv.reshape(batch, n, d), r["sysvar"].reshape(batch, n, d)
            assert rel_field(a[..., :M], b[..., :M]) <= RTOL_CHAIN_V
            assert rel_field(a[..., M:], b[..., M:]) <= RTOL_CHAIN
            host.sysvar, host.syslod = sv, sl


def test_full_size_mesh_batch_order_invariance():
    """BASELINE mesh size (10 001 nodes) through a size-independent property: a conductor's result does
    not depend on its slot in the batch or on its neighbours -- K_P + step on a batch and on its
    reversal agree bit for bit, slot by slot (also covers a ragged last tile: 37 conductors)."""
    from opensc2_b200.batch import BatchedStep

    n, batch, dt = 10001, 37, 0.1
    topo, model, tab, _a, sweep = _setup(True, 17, batch, seed=31)
    arr = consistent_initial_state(topo, model, tab, n, batch, seed=31, dt=dt)
    sweep = synthetic_sweep(topo, model, n, batch, seed=31)
    outs = []
    for order in (np.arange(batch), np.arange(batch)[::-1]):
        import dataclasses
        sub = type(arr)(**{f.name: (getattr(arr, f.name)[order] if isinstance(getattr(arr, f.name), np.ndarray) else getattr(arr, f.name))
                           for f in dataclasses.fields(arr)})
        sw = type(sweep)(**{f.name: (getattr(sweep, f.name)[order] if getattr(sweep, f.name) is not None else None)
                            for f in dataclasses.fields(sweep)})
        with BatchedStep(topo, batch, n) as bs:
            bs.set_model(model, [tab])
            bs.upload_sweep(sw)
            bs.upload_state(sub.sysvar, sub.syslod)
            bs.upload_geometry(sub)
            for k in range(1, 3):
                bs.advance(9.9 + k * dt, dt, k)
            bs.synchronize()
            sv, _sl, st = bs.download_state(want_syslod=False)
            diag = bs.download_diagnostics(want_k123=False)
        assert np.all(st == 0) and np.all(np.isfinite(sv))
        outs.append((sv[np.argsort(order)], diag["norm_change"][np.argsort(order)]))
    assert_bits(outs[0][0], outs[1][0], "state")
    assert_bits(outs[0][1], outs[1][1], "norm_change")


def test_case3_flavoured_model_radiative_exchange():
    """d = 7: radiation between two jackets (HTC_choice 3, nodal HTC, explicit source on both solids), a
    constant contact, convection to the environment; K_P vs oracle (sources bit-exact) and two chained
    steps."""
    from opensc2_b200.batch import BatchedStep
    from opensc2_b200.model import case3_mini_model
    from opensc2_b200.problem import case3_mini_topology
    from oracle import oracle as O

    n, batch, dt = 21, 5, 0.1
    topo, model, tab = case3_mini_topology(), case3_mini_model(), synthetic_helium_table()
    arr = consistent_initial_state(topo, model, tab, n, batch, seed=6, dt=dt)
    sv0 = arr.sysvar.reshape(batch, n, topo.ndf)
    sv0[:, :, topo.ndf - 3:] += np.array([10.0, 40.0, 60.0])[None, None, :]
    rng = np.random.default_rng(2)
    arr.peri_ss_nodal = rng.uniform(0.05, 0.06, (batch, len(topo.ss), 1)) * np.ones(n)
    sweep = synthetic_sweep(topo, model, n, batch, seed=6, with_tcs=False)
    host = arr.copy()
    d, M = topo.ndf, topo.n_ch
    with BatchedStep(topo, batch, n) as bs:
        bs.set_model(model, [tab])
        bs.upload_sweep(sweep)
        bs.upload_state(arr.sysvar, arr.syslod)
        bs.upload_geometry(arr)
        bs.set_capture(True)
        for k in range(1, 3):
            t = 11.9 + k * dt
            bs.advance(t, dt, k)
            bs.synchronize()
            ins = bs.download_step_inputs()
            sv, sl, st = bs.download_state()
            assert np.all(st == 0)
            ref_in = O.operating_conditions(topo, model, [tab], sweep, host.sysvar, n, t, peri_ss_nodal=arr.peri_ss_nodal)
            assert_bits(ins["sol_q"], ref_in["sol_q"], "Q1/Q2 with radiative exchange")
            assert np.abs(ref_in["sol_q"][:, :, 1:3]).max() > 0
            for key in ins:
                assert rel(ins[key], ref_in[key]) <= RTOL_FIT, f"{key}: {rel(ins[key], ref_in[key]):.3e}"
            for key, v in ref_in.items():
                setattr(host, key, v)
            host.num_step, host.dt = k, dt
            r = O.step(topo, host)
            a, b = sv.reshape(batch, n, d), r["sysvar"].reshape(batch, n, d)
            assert rel_field(a[..., :M], b[..., :M]) <= RTOL_CHAIN_V
            assert rel_field(a[..., M:], b[..., M:]) <= RTOL_CHAIN
            host.sysvar, host.syslod = sv, sl


def test_stack_component_solid_on_the_device():
    """SOLID_STACK (StackComponent: Ag / Cu / Hastelloy C276 / REBCO / Sn60Pb40 layers homogenised by
    thickness, stack_component.py:398-476): K_P against the oracle and against the values the reference's
    own StackComponent methods gave for the same temperatures and fields (tests/golden/props_stack.npz)."""
    from helpers import stack_case
    from opensc2_b200.batch import BatchedStep
    from oracle import oracle as O

    topo, model, tab, arr, sweep, g = stack_case()
    n, batch = arr.n_nodes, arr.sysvar.shape[0]
    with BatchedStep(topo, batch, n) as bs:
        bs.set_model(model, [tab])
        bs.upload_sweep(sweep)
        bs.upload_state(arr.sysvar, arr.syslod)
        bs.upload_geometry(arr)
        bs.eval_operating_conditions(5.0)
        bs.synchronize()
        out = bs.download_step_inputs()
    ref = O.operating_conditions(topo, model, [tab], sweep, arr.sysvar, n, 5.0)
    assert rel(out["sol_gauss"], ref["sol_gauss"]) <= RTOL_FIT
    for q, key in enumerate(("stack_rho", "stack_cp", "stack_k")):
        for e in range(n - 1):
            assert np.max(np.abs(out["sol_gauss"][:, q, 0, e] - g[key]) / np.abs(g[key])) <= RTOL_FIT, key


@pytest.mark.parametrize("laws", [(110, 121), (111, 112), (204, 206), (205, 209), (113, 114), (115, 118), (200, 201), (202, 203),
                                  (207, 210), (211, 113), (100, 101), (102, 103), (104, 105), (106, 107), (108, 109), (120, 106), (119, 120)])
def test_more_friction_laws_on_the_device(laws):
    """IFRICTION 110 (Blasius), 111 (Bhatti-Shah), 112 (LHC recipe), 121 (Haaland), 204 / 205 (Katheder bundle),
    206 / 209 (Darcy-Forchheimer bundle) in K_P vs the oracle (which
    is pinned to the reference's Channel.eval_friction_factor, tests/test_props_oracle_golden.py)."""
    import dataclasses
    from opensc2_b200.batch import BatchedStep
    from oracle import oracle as O

    n, batch = 33, 21
    topo, model, tab, arr, sweep = _setup(True, n, batch)
    model = dataclasses.replace(model, ch_ifriction=laws, ch_roughness=(2.0e-6, 2.0e-6), ch_void_fraction=(0.33, 0.36),
                                ch_side_ratio=(0.4, 0.7))
    with BatchedStep(topo, batch, n) as bs:
        bs.set_model(model, [tab])
        bs.upload_sweep(sweep)
        bs.upload_state(arr.sysvar, arr.syslod)
        bs.upload_geometry(arr)
        bs.eval_operating_conditions(5.0)
        bs.synchronize()
        out = bs.download_step_inputs()
    ref = O.operating_conditions(topo, model, [tab], sweep, arr.sysvar, n, 5.0)
    assert np.all(ref["fl_gauss"][:, 8] > 0.0)
    assert rel(out["fl_gauss"][:, 8], ref["fl_gauss"][:, 8]) <= RTOL_FIT
    assert rel(out["fl_gauss"], ref["fl_gauss"]) <= RTOL_FIT


@pytest.mark.parametrize("sc", ["nbti", "mgb2"])
def test_niobium_titanium_strand_on_the_device(sc):
    """StrandMixedComponent with NbTi (Tc from critical_temperature_nbti, cp from isobaric_specific_heat_nbti with
    the current-sharing temperature of the sweep) or MgB2 (conductivity of temperature and field, cp of
    temperature) as superconductor: K_P vs the oracle, which is pinned to the reference fits
    (tests/test_props_oracle_golden.py)."""
    import dataclasses
    from opensc2_b200 import model as Mdl
    from opensc2_b200.batch import BatchedStep
    from oracle import oracle as O

    n, batch = 33, 19
    topo, model, tab, arr, sweep = _setup(True, n, batch)
    strand = Mdl.SolidModel(Mdl.SOLID_STRAND_MIXED, (Mdl.MAT_CU_NIST, Mdl.MAT_NBTI if sc == "nbti" else Mdl.MAT_MGB2), (1.6, 1.0), 2.6, rrr=120.0,
                            tc0m=9.03, bc20m=14.5, heated=True)
    model = dataclasses.replace(model, solids=(strand,) + model.solids[1:])
    with BatchedStep(topo, batch, n) as bs:
        bs.set_model(model, [tab])
        bs.upload_sweep(sweep)
        bs.upload_state(arr.sysvar, arr.syslod)
        bs.upload_geometry(arr)
        bs.eval_operating_conditions(12.0)
        bs.synchronize()
        out = bs.download_step_inputs()
    ref = O.operating_conditions(topo, model, [tab], sweep, arr.sysvar, n, 12.0)
    assert rel(out["sol_gauss"], ref["sol_gauss"]) <= RTOL_FIT
    for key in ("htc_fs", "htc_ss", "sol_q"):
        assert rel(out[key], ref[key]) <= RTOL_FIT, key


@pytest.mark.parametrize("corrs", [(119, 211), (120, 121)])
def test_more_steady_state_htc_correlations_on_the_device(corrs):
    """Flag_htc_steady_corr 119 / 120 / 121 / 211 feeding the channel-solid and channel-channel heat-transfer
    models (HTC_choice 2) in K_P vs the oracle."""
    import dataclasses
    from opensc2_b200.batch import BatchedStep
    from oracle import oracle as O

    n, batch = 33, 21
    topo, model, tab, arr, sweep = _setup(True, n, batch)
    model = dataclasses.replace(model, ch_htc_corr=corrs)
    with BatchedStep(topo, batch, n) as bs:
        bs.set_model(model, [tab])
        bs.upload_sweep(sweep)
        bs.upload_state(arr.sysvar, arr.syslod)
        bs.upload_geometry(arr)
        bs.eval_operating_conditions(12.0)
        bs.synchronize()
        out = bs.download_step_inputs()
    ref = O.operating_conditions(topo, model, [tab], sweep, arr.sysvar, n, 12.0)
    for key in ("htc_fs", "htc_ff", "htc_ss", "fl_gauss", "sol_gauss"):
        assert rel(out[key], ref[key]) <= RTOL_FIT, f"{key}: {rel(out[key], ref[key]):.3e}"
    if 211 in corrs:      # (the other three only move the lower limit, which this flow does not reach)
        base = O.operating_conditions(topo, dataclasses.replace(model, ch_htc_corr=(1, 1)), [tab], sweep, arr.sysvar, n, 12.0)
        assert np.max(np.abs(base["htc_fs"] - ref["htc_fs"])) > 0.0


@pytest.mark.parametrize("colebrook", [False, True])
def test_nodal_coolant_pass_matches_the_host_writer(colebrook):
    """osc2_eval_nodal_properties (post-step density, mass flow rate, the nodal property columns and the nodal
    friction factor, coolant.py:134-263 with nodal = True) against opensc2_b200.output.channel_table, the host code
    pinned byte for byte to files written by the reference's save_properties (tests/test_output_writer.py): table
    columns, Reynolds, Gruneisen and mass flow rate bit-exact, the friction factor within RTOL_FIT."""
    import dataclasses

    from opensc2_b200.batch import BatchedStep
    from opensc2_b200.output import channel_table

    n, batch = 41, 9
    topo, model, tab, arr, sweep = _setup(True, n, batch, seed=11)
    if colebrook:        # IFRICTION 122: the iteration count is the maximum over the nodes of one channel
        model = dataclasses.replace(model, ch_ifriction=(122,) * topo.n_ch, ch_roughness=(2.0e-6,) * topo.n_ch)
    with BatchedStep(topo, batch, n) as bs:
        bs.set_model(model, [tab])
        bs.upload_sweep(sweep)
        bs.upload_state(arr.sysvar, arr.syslod)
        out = bs.nodal_properties()
    z = np.linspace(0.0, 10.0, n)
    x = arr.sysvar.reshape(batch, n, topo.ndf)
    M = topo.n_ch
    for b in range(batch):
        for j in range(M):
            ref = channel_table(topo, model, tab, j, z, x[b, :, j].copy(), x[b, :, M + j].copy(), x[b, :, 2 * M + j].copy())
            assert_bits(out["total_density"][b, j], ref[:, 3], "total_density")
            for i, name in enumerate(BatchedStep.NODAL_COLUMNS[1:13]):
                assert_bits(out[name][b, j], ref[:, 5 + i], name)
            assert rel(out["friction_factor"][b, j], ref[:, 17]) <= RTOL_FIT
