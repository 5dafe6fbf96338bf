"""The CPU oracle reproduces the reference step() fixtures bit for bit.

Fixtures: tests/golden/step_*.npz, written by tests/golden/make_golden.py from
the unmodified reference (transient_solution_functions.py:186-599).
"""
import numpy as np
import pytest

from cases import STEP_CASES
from helpers import assert_bits, load_golden
from oracle import oracle as O


@pytest.mark.parametrize("name", sorted(STEP_CASES))
def test_step_oracle_matches_reference_fixture(name):
    topo, arr, ref = load_golden(name)
    out = O.step(topo, arr, want_system=True)
    assert out["status"] == 0
    for key in ("sysmat", "known"):
        if key in ref:
            assert_bits(out[key], ref[key], f"{name}:{key}")
    for key in ("sysvar", "syslod", "k123", "norm_solution", "norm_change", "eqteig"):
        assert_bits(out[key], ref[key], f"{name}:{key}")


def test_fixture_inputs_regenerate_from_seed():
    """The committed inputs are what the synthetic generator produces today."""
    from cases import make_case

    for name in ("case1_pdrop_n21", "case3_backward_n15"):
        _t, arr, _ref = load_golden(name)
        _t2, arr2 = make_case(name)
        assert_bits(arr2.fl_gauss, arr.fl_gauss, "fl_gauss")
        assert_bits(arr2.sysvar, arr.sysvar, "sysvar")


def test_pairwise_sum_is_numpy_sum():
    rng = np.random.default_rng(1)
    for n in (1, 5, 7, 8, 9, 31, 51, 127, 128, 129, 255, 1000, 10001):
        a = rng.standard_normal(n) * 10.0 ** rng.integers(-6, 6, n)
        assert O.pairwise_sum(a) == np.sum(a)
        b = np.zeros(n * 8)
        b[3::8] = a
        assert O.pairwise_sum(b[3:], 8) == np.sum(b[3::8])


def test_batched_oracle_equals_single():
    from opensc2_b200.problem import case1_topology, synthetic_step_arrays, take

    topo = case1_topology(2)
    arr = synthetic_step_arrays(topo, 9, batch=5)
    out = O.step(topo, arr, n_threads=3)
    for b in range(5):
        one = O.step(topo, take(arr, b))
        assert_bits(out["sysvar"][b], one["sysvar"], f"b={b}")
        assert_bits(out["norm_change"][b], one["norm_change"], f"b={b}")


def test_singular_pivot_reported():
    """Reference: `ValueError: Matrix is singular at line 8` (checked live in the
    build container); the oracle reports -(8+1) for that conductor only."""
    from helpers import singular_case

    topo, arr = singular_case()
    out = O.step(topo, arr)
    assert list(out["status"]) == [0, -9]
