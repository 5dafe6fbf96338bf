"""Kernel LOGIC on CPU: the CUDA step kernels, compiled with g++ against a
thread emulator (tests/cuda_emu), reproduce the oracle bit for bit.  This does
not replace the `-m gpu` parity tests; it catches index and ordering errors
where no GPU is available."""
import numpy as np
import pytest

from helpers import assert_bits, load_golden
from opensc2_b200.problem import (case1_topology, case2_topology, case3_topology,
                                  synthetic_step_arrays)
from oracle import oracle as O

emu = pytest.importorskip("cuda_emu.emu")

KEYS = ("sysvar", "syslod", "k123", "norm_solution", "norm_change", "eqteig")


def _batched(arr):
    import dataclasses
    kw = {}
    for f in dataclasses.fields(arr):
        v = getattr(arr, f.name)
        kw[f.name] = v[None].copy() if isinstance(v, np.ndarray) else v
    return type(arr)(**kw)


@pytest.mark.parametrize("name", ["case1_pdrop_n21", "case1_pinl_vout_first_step_n17",
                                  "case1_vinl_pout_cn_refined_n12", "case1_min_mesh_n4",
                                  "case3_backward_n15", "case3_rectangular_env_n9"])
@pytest.mark.parametrize("fast,split", [(False, 0), (True, 0)])
def test_emulated_kernels_match_reference_fixture(name, fast, split):
    topo, arr, ref = load_golden(name)
    out = emu.emu_step(topo, _batched(arr), capture=True, fast=fast, split=split)
    assert out["status"][0] == 0
    assert_bits(out["sysmat"][0], ref["sysmat"], f"{name}:sysmat")
    assert_bits(out["known"][0], ref["known"], f"{name}:known")
    for key in KEYS:
        assert_bits(out[key][0], ref[key], f"{name}:{key}")


@pytest.mark.parametrize("n", [129, 201, 300, 1100])
def test_emulated_norm_tree_matches_numpy_pairwise(n):
    """Leaf-parallel norms (K_N1/K_N2) replay numpy's pairwise-sum tree for N > 128."""
    topo = case1_topology(1)
    arr = synthetic_step_arrays(topo, n, batch=2, seed=n)
    out = emu.emu_step(topo, arr)
    ref = O.step(topo, arr)
    for key in KEYS:
        assert_bits(out[key], ref[key], key)


@pytest.mark.parametrize("n,batch,num_step", [(4, 2, 1), (6, 3, 3)])
def test_emulated_d64_register_tiled_sweeps(n, batch, num_step):
    """step_kernels_d64.cuh (256 threads per conductor, rows in registers, warp-per-conductor back
    substitution): SYSMAT, Known and every output equal the oracle's bits; the generic shared-memory
    sweeps (OSC2_D64=0 in the library) are covered by test_emulated_case2_sixty_four_lanes' history
    and by the GPU tests."""
    topo = case2_topology()
    arr = synthetic_step_arrays(topo, n, batch=batch, seed=60 + n, num_step=num_step, exact_squares=False)
    out = emu.emu_step(topo, arr, capture=True)
    ref = O.step(topo, arr, want_system=True)
    assert np.all(out["status"] == 0)
    assert_bits(out["sysmat"], ref["sysmat"], "sysmat")
    assert_bits(out["known"], ref["known"], "known")
    for key in KEYS:
        assert_bits(out[key], ref[key], key)


def test_emulated_d64_singular_status():
    from helpers import case2_with_isolated_massless_solid

    topo, arr, row = case2_with_isolated_massless_solid(5, 2, bad=1)
    ref = O.step(topo, arr)
    assert list(np.atleast_1d(ref["status"])) == [0, -(row + 1)]
    out = emu.emu_step(topo, arr)
    assert list(out["status"]) == [0, -(row + 1)]
    assert_bits(out["sysvar"][0], ref["sysvar"][0], "healthy conductor next to the singular one")


def test_emulated_d64_flag_synchronised_rounds():
    """Opt-in OSC2_D64_SYNC=flags (per-row shared-memory flags instead of a CTA barrier per in-node pivot
    round): same bits.  The switch is read once per process, hence the child."""
    import os, subprocess, sys, textwrap
    here = os.path.dirname(os.path.abspath(__file__))
    code = textwrap.dedent(f"""
        import sys, numpy as np
        sys.path[:0] = [{here!r}, {os.path.dirname(here)!r}]
        from cuda_emu import emu
        from opensc2_b200.problem import case2_topology, synthetic_step_arrays
        from oracle import oracle as O
        topo = case2_topology()
        arr = synthetic_step_arrays(topo, 5, batch=2, seed=31, num_step=2, exact_squares=False)
        out = emu.emu_step(topo, arr, capture=True)
        ref = O.step(topo, arr, want_system=True)
        for key in ("sysmat", "known", "sysvar", "syslod", "norm_solution", "norm_change", "eqteig"):
            assert np.array_equal(out[key], ref[key]), key
        assert np.all(out["status"] == 0)
        print("bits-equal")
    """)
    res = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, OSC2_D64_SYNC="flags"),
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "bits-equal" in res.stdout, res.stdout + res.stderr


@pytest.mark.parametrize("intial,n,batch,num_step", [(1, 9, 8, 2), (2, 7, 6, 2), (3, 23, 5, 3), (1, 4, 4, 1), (3, 5, 13, 1), (2, 40, 1, 2)])
def test_emulated_tile_sweep_batches(intial, n, batch, num_step):
    """step_kernels_tile.cuh (tiled records through bulk copies, node-pipelined rounds): full and ragged tiles, every
    boundary-condition mode, first step, the minimum mesh (3 elements in a ring of 3 stages), against the oracle."""
    topo = case1_topology(intial)
    arr = synthetic_step_arrays(topo, n, batch=batch, seed=40 + n, num_step=num_step)
    out = emu.emu_step(topo, arr, fast=True)
    ref = O.step(topo, arr)
    assert np.all(out["status"] == 0)
    for key in KEYS:
        assert_bits(out[key], ref[key], key)


def test_emulated_tile_sweep_reports_a_singular_conductor():
    topo = case1_topology(1)
    arr = synthetic_step_arrays(topo, 9, batch=6, seed=3)
    # conductor 4: the jacket without mass term and without thermal contact -- its Neumann diffusion rows
    # [1 -1], [-1/2 1 -1/2], ..., [-1 1] eliminate to an exact zero pivot in the last node (tsf:763-786)
    arr.sol_gauss[4, 1, 1] = 0.0
    arr.sol_gauss[4, 2, 1] = 100.0
    arr.htc_fs[4, 1], arr.htc_ss[4], arr.htc_es[4] = 0.0, 0.0, 0.0
    arr.sol_q[4, :, 1] = 0.0
    out = emu.emu_step(topo, arr, fast=True)
    ref = O.step(topo, arr)
    assert list(out["status"]) == list(ref["status"]) and ref["status"][4] != 0
    ok = np.array([0, 1, 2, 3, 5])
    assert_bits(out["sysvar"][ok], ref["sysvar"][ok], "the conductors next to the singular one")


@pytest.mark.parametrize("switch", ["OSC2_FWD_SPLIT", "OSC2_FUSED"])
def test_emulated_forward_sweep_experiments_stay_bit_exact(switch):
    """experiments/step_kernels_split_roles.cuh (assembler + eliminator warps per tile, hand-over through mbarriers) and
    experiments/step_kernels_fused.cuh (K_G as producer warps inside K_F): not in the product build (both measured slower,
    DESIGN.md 5), compiled into the emulator library so that their logic stays verified.  The switch is read once per
    process, hence the child."""
    import os, subprocess, sys, textwrap
    here = os.path.dirname(os.path.abspath(__file__))
    code = textwrap.dedent(f"""
        import sys, numpy as np
        sys.path[:0] = [{here!r}, {os.path.dirname(here)!r}]
        from cuda_emu import emu
        from opensc2_b200.problem import case1_topology, synthetic_step_arrays
        from oracle import oracle as O
        for intial, n, batch, ns in ((1, 9, 8, 2), (2, 23, 37, 3), (3, 5, 29, 1)):
            topo = case1_topology(intial)
            arr = synthetic_step_arrays(topo, n, batch=batch, seed=40 + n, num_step=ns)
            out = emu.emu_step(topo, arr, fast=True)
            ref = O.step(topo, arr)
            for key in ("sysvar", "syslod", "k123", "norm_solution", "norm_change", "eqteig"):
                assert np.array_equal(out[key], ref[key]), key
            assert np.all(out["status"] == 0)
        print("bits-equal", emu.last_launches)
    """)
    res = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **{switch: "1"}), capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "bits-equal" in res.stdout, res.stdout + res.stderr
    if switch == "OSC2_FUSED":
        assert res.stdout.split()[-1] == "4"          # K_GF + K_B + two norm kernels: K_G's launch is gone

