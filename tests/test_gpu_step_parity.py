"""GPU parity: the sm_100a kernels, called through the C ABI, against the
reference fixtures and the CPU oracle.  Bit-exact (FP64, -fmad=false)."""
import dataclasses

import numpy as np
import pytest

from cases import STEP_CASES, WITH_SYSTEM
from helpers import assert_bits, load_golden
from opensc2_b200.problem import (case1_topology, case2_topology, case3_topology,
                                  synthetic_step_arrays)

pytestmark = pytest.mark.gpu

KEYS = ("sysvar", "syslod", "k123", "norm_solution", "norm_change", "eqteig")


def _batched(arr):
    kw = {}
    for f in dataclasses.fields(arr):
        v = getattr(arr, f.name)
        kw[f.name] = v[None].copy() if isinstance(v, np.ndarray) else v
    return type(arr)(**kw)


@pytest.mark.parametrize("name", sorted(STEP_CASES))
def test_reference_fixture_bit_exact(name):
    from opensc2_b200.batch import BatchedStep

    topo, arr, ref = load_golden(name)
    cap = name in WITH_SYSTEM
    with BatchedStep(topo, 1, arr.n_nodes) as bs:
        out = bs.run(_batched(arr), capture=cap)
    assert out["status"][0] == 0
    if cap:
        assert_bits(out["sysmat"][0], ref["sysmat"], f"{name}:sysmat")
        assert_bits(out["known"][0], ref["known"], f"{name}:known")
    for key in KEYS:
        assert_bits(out[key][0], ref[key], f"{name}:{key}")


@pytest.mark.parametrize("label,topo,n,batch", [
    ("case1", case1_topology(1), 65, 37),
    ("case1_intial3", case1_topology(3), 130, 64),
    ("case3", case3_topology(), 40, 9),
    ("case2", case2_topology(), 7, 3),
])
def test_batch_matches_oracle(label, topo, n, batch):
    from opensc2_b200.batch import BatchedStep
    from oracle import oracle as O

    arr = synthetic_step_arrays(topo, n, batch=batch, seed=17, exact_squares=False)
    with BatchedStep(topo, batch, n) as bs:
        out = bs.run(arr)
    ref = O.step(topo, arr)
    assert np.all(out["status"] == 0)
    for key in KEYS:
        assert_bits(out[key], ref[key], f"{label}:{key}")


def test_chained_steps_match_oracle():
    """Five consecutive steps (state and SYSLOD history carried on the device)."""
    from opensc2_b200.batch import BatchedStep
    from oracle import oracle as O

    topo = case1_topology(1)
    n, batch = 41, 12
    arr = synthetic_step_arrays(topo, n, batch=batch, seed=23, num_step=1)
    host = arr.copy()
    with BatchedStep(topo, batch, n) as bs:
        bs.upload_state(arr.sysvar, arr.syslod)
        bs.upload_step_inputs(arr)
        for k in range(1, 6):
            bs.step(arr.dt, k)
            host.num_step = k
            ref = O.step(topo, host)
            host.sysvar, host.syslod = ref["sysvar"], ref["syslod"]
        bs.synchronize()
        sv, sl, st = bs.download_state()
    assert np.all(st == 0)
    assert_bits(sv, host.sysvar, "sysvar after 5 steps")
    assert_bits(sl, host.syslod, "syslod after 5 steps")


def test_full_size_mesh_linearity_property():
    """10k-node mesh (BASELINE size per conductor): solution of the captured system
    satisfies SYSMAT x = Known to FP64 round-off (size-independent residual check)."""
    from opensc2_b200.batch import BatchedStep

    topo = case1_topology(1)
    n, batch = 10001, 4
    arr = synthetic_step_arrays(topo, n, batch=batch, seed=3)
    with BatchedStep(topo, batch, n) as bs:
        out = bs.run(arr, capture=True)
    d = topo.ndf
    total = d * n
    main = 2 * d - 1
    x = out["sysvar"]
    for b in range(batch):
        band, known = out["sysmat"][b], out["known"][b]
        ax = np.zeros(total)
        for k in range(4 * d - 1):
            off = k - main
            lo, hi = max(0, -off), min(total, total - off)
            ax[lo:hi] += band[k, lo:hi] * x[b, lo + off:hi + off]
        scale = np.abs(band).T @ np.ones(4 * d - 1) * np.max(np.abs(x[b]))
        assert np.max(np.abs(ax - known) / scale) < 1e-9


@pytest.mark.parametrize("batch", [8, 29])
def test_full_size_mesh_matches_the_oracle_bit_for_bit(batch):
    """BASELINE mesh size per conductor (10 001 nodes) through the production sweeps (split-role K_F, tiled K_B),
    a full and a ragged set of tiles, two chained steps: every output equals the oracle's bits."""
    from opensc2_b200.batch import BatchedStep
    from oracle import oracle as O

    topo = case1_topology(1)
    n = 10001
    arr = synthetic_step_arrays(topo, n, batch=batch, seed=7, num_step=1)
    host = arr.copy()
    with BatchedStep(topo, batch, n) as bs:
        bs.upload_state(arr.sysvar, arr.syslod)
        bs.upload_step_inputs(arr)
        for k in (1, 2):
            bs.step(arr.dt, k)
            host.num_step = k
            ref = O.step(topo, host, n_threads=8) if "n_threads" in O.step.__code__.co_varnames else O.step(topo, host)
            host.sysvar, host.syslod = ref["sysvar"], ref["syslod"]
        bs.synchronize()
        sv, sl, st = bs.download_state()
        diag = bs.download_diagnostics()
    assert np.all(st == 0)
    assert_bits(sv, host.sysvar, "sysvar, 10 001 nodes")
    assert_bits(sl, host.syslod, "syslod, 10 001 nodes")
    for key in ("norm_solution", "norm_change", "eqteig"):
        assert_bits(diag[key], ref[key], key)


def test_too_few_nodes_is_rejected():
    from opensc2_b200.batch import BatchedStep

    with pytest.raises(ValueError):
        BatchedStep(case1_topology(1), 2, 3)


def test_singular_system_raises_like_reference():
    """Pure-solid Neumann diffusion without a mass term: the reference raises
    `ValueError: Matrix is singular at line 8` (tsf:763-786); here status = -(8+1)."""
    from opensc2_b200 import _lib
    from opensc2_b200.batch import BatchedStep
    from helpers import singular_case

    topo, arr = singular_case()
    with BatchedStep(topo, 2, arr.n_nodes) as bs:
        bs.upload_state(arr.sysvar, arr.syslod)
        bs.upload_step_inputs(arr)
        bs.step(arr.dt, arr.num_step)
        with pytest.raises(_lib.SingularMatrixError):
            bs.synchronize()
        _sv, _sl, st = bs.download_state(want_syslod=False)
    assert st[0] == 0 and st[1] == -9


def test_value_outside_the_exact_division_range_is_reported_not_returned():
    """A non-zero matrix entry below 2^-900 cannot go through the shared-divisor division
    bit-exactly; the sweep must say so for that conductor (OSC2_STATUS_RANGE) and leave the
    rest of the batch bit-exact."""
    from opensc2_b200._abi import OSC2_STATUS_RANGE
    from opensc2_b200.batch import BatchedStep
    from oracle import oracle as O

    topo = case1_topology(1)
    n, batch = 33, 8
    arr = synthetic_step_arrays(topo, n, batch=batch, seed=5, exact_squares=False)
    from opensc2_b200 import _lib

    arr.htc_fs[3, 0, 10] = 1.0e-290          # times a perimeter and dz/3: far below 2^-900 = 1.2e-271
    with BatchedStep(topo, batch, n) as bs:
        bs.upload_state(arr.sysvar, arr.syslod)
        bs.upload_step_inputs(arr)
        bs.step(arr.dt, arr.num_step)
        with pytest.raises(_lib.Osc2Error) as err:
            bs.synchronize()
        assert err.value.code == -6
        sv, sl, st = bs.download_state()
    ok = np.arange(batch) != 3
    assert st[3] == OSC2_STATUS_RANGE and np.all(st[ok] == 0)
    ref = O.step(topo, arr)
    assert_bits(sv[ok], ref["sysvar"][ok], "range:sysvar")
    assert_bits(sl[ok], ref["syslod"][ok], "range:syslod")


def test_shared_divisor_division_is_ieee():
    """csrc/osc2_div.cuh returns exactly a / b (incl. signed zeros, tiny/huge operands)."""
    import ctypes as C

    from opensc2_b200 import _lib

    L = _lib.load()
    bad = C.c_ulonglong(123)
    _lib.check(L.osc2_selftest_division(0, 400_000_000, 20261019, C.byref(bad)))
    assert bad.value == 0


@pytest.mark.parametrize("n,batch,num_step,d64", [(151, 5, 2, "1"), (9, 3, 1, "1"), (9, 3, 1, "0"), (31, 4, 2, "flags")])
def test_case2_sized_sweeps_match_oracle_with_system(n, batch, num_step, d64, monkeypatch):
    """NODOFS = 64 at the shipped CASE_2 mesh size: the register-tiled sweeps (step_kernels_d64.cuh,
    default; also with the opt-in per-row flags, OSC2_D64_SYNC=flags) and the generic shared-memory sweeps (OSC2_D64=0) reproduce the oracle's SYSMAT,
    Known and outputs bit for bit."""
    import subprocess, sys, os, textwrap
    # the switch is read once per process: run the comparison in a child
    code = textwrap.dedent(f"""
        import sys, numpy as np
        sys.path[:0] = [{os.path.dirname(os.path.abspath(__file__))!r}, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r}]
        from opensc2_b200.batch import BatchedStep
        from opensc2_b200.problem import case2_topology, synthetic_step_arrays
        from oracle import oracle as O
        topo = case2_topology()
        arr = synthetic_step_arrays(topo, {n}, batch={batch}, seed=77, num_step={num_step}, exact_squares=False)
        with BatchedStep(topo, {batch}, {n}) as bs:
            out = bs.run(arr, capture=True)
        ref = O.step(topo, arr, want_system=True)
        assert np.all(out["status"] == 0)
        for key in ("sysmat", "known", "sysvar", "syslod", "k123", "norm_solution", "norm_change", "eqteig"):
            assert np.array_equal(out[key], ref[key]), key
        print("bits-equal")
    """)
    env = dict(os.environ, OSC2_D64=d64) if d64 != "flags" else dict(os.environ, OSC2_D64="1", OSC2_D64_SYNC="flags")
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "bits-equal" in res.stdout, res.stdout + res.stderr


def test_case2_singular_conductor_is_reported_by_the_d64_sweep():
    """NODOFS = 64: an exact zero pivot in one conductor gives status -(row+1) (`Matrix is singular at
    line row`, tsf:763-786) while its neighbours in the batch stay bit-exact."""
    from helpers import case2_with_isolated_massless_solid
    from opensc2_b200 import _lib
    from opensc2_b200.batch import BatchedStep
    from oracle import oracle as O

    topo, arr, row = case2_with_isolated_massless_solid(7, 3, bad=1)
    ref = O.step(topo, arr)
    assert list(ref["status"]) == [0, -(row + 1), 0]
    with BatchedStep(topo, 3, 7) as bs:
        bs.upload_state(arr.sysvar, arr.syslod)
        bs.upload_step_inputs(arr)
        bs.step(arr.dt, arr.num_step)
        with pytest.raises(_lib.SingularMatrixError):
            bs.synchronize()
        sv, _sl, st = bs.download_state(want_syslod=False)
    assert list(st) == [0, -(row + 1), 0]
    assert_bits(sv[0], ref["sysvar"][0], "conductor 0")
    assert_bits(sv[2], ref["sysvar"][2], "conductor 2")


def test_case1_and_case2_handles_side_by_side():
    """The mixed sweep of BASELINE configs[4] keeps one handle per topology on a GPU: a CASE_1 batch (register-
    resident warp kernels) and a CASE_2 batch (CTA-per-conductor d = 64 sweeps) alive at the same time, stepped
    alternately for three steps, each bit-identical to the oracle run on its own."""
    from opensc2_b200.batch import BatchedStep
    from oracle import oracle as O

    t1, t2 = case1_topology(2), case2_topology()
    a1 = synthetic_step_arrays(t1, 201, batch=48, seed=61, num_step=1, exact_squares=False)
    a2 = synthetic_step_arrays(t2, 31, batch=6, seed=62, num_step=1, exact_squares=False)
    h1, h2 = a1.copy(), a2.copy()
    with BatchedStep(t1, 48, 201) as b1, BatchedStep(t2, 6, 31) as b2:
        b1.upload_state(a1.sysvar, a1.syslod); b1.upload_step_inputs(a1)
        b2.upload_state(a2.sysvar, a2.syslod); b2.upload_step_inputs(a2)
        for k in range(1, 4):
            b1.step(a1.dt, k)
            b2.step(a2.dt, k)
            for host, topo in ((h1, t1), (h2, t2)):
                host.num_step = k
                ref = O.step(topo, host)
                host.sysvar, host.syslod = ref["sysvar"], ref["syslod"]
        b1.synchronize(); b2.synchronize()
        sv1, sl1, st1 = b1.download_state()
        sv2, sl2, st2 = b2.download_state()
    assert np.all(st1 == 0) and np.all(st2 == 0)
    assert_bits(sv1, h1.sysvar, "CASE_1 sysvar"); assert_bits(sl1, h1.syslod, "CASE_1 syslod")
    assert_bits(sv2, h2.sysvar, "CASE_2 sysvar"); assert_bits(sl2, h2.syslod, "CASE_2 syslod")
