"""Host logic of the drop-in `step(conductor, environment, qsource, num_step)` shim
(opensc2_b200/step.py): flatten -> device call -> scatter, on duck-typed conductors
that carry the attributes the reference reads (tests/fake_conductor.py)."""
import dataclasses

import numpy as np
import pytest

from fake_conductor import build_conductor
from helpers import assert_bits, load_golden
from opensc2_b200 import step as shim

CASES = ["case1_pdrop_n21", "case1_pinl_vout_first_step_n17", "case3_rectangular_env_n9", "case2_n6"]


@pytest.mark.parametrize("name", CASES)
def test_pack_round_trip(name):
    topo, arr, _ref = load_golden(name)
    cond, env, qsource = build_conductor(topo, arr)
    topo2 = shim.topology_of(cond)
    assert topo2 == topo
    arr2 = shim.pack(cond, env, qsource, arr.num_step, topo2)
    for f in dataclasses.fields(arr):
        a, b = getattr(arr, f.name), getattr(arr2, f.name)
        if isinstance(a, np.ndarray):
            if f.name == "htc_es" and not (topo.is_rectangular and any(topo.es_choice2)):
                a, b = a[:, 0], b[:, 0]          # bottom/top only exist for rectangular ducts
            assert_bits(b, a, f.name)
        else:
            assert a == b, f.name


@pytest.mark.parametrize("name", CASES[:2])
def test_unpack_writes_what_the_reference_writes(name):
    """Feed the fixture's reference outputs through unpack(); every attribute the
    reference step() writes must appear with the reference values."""
    topo, arr, ref = load_golden(name)
    cond, env, qsource = build_conductor(topo, arr)
    out = {k: v[None] for k, v in ref.items()}
    shim.unpack(cond, out, 0, topo)
    M, d = topo.n_ch, topo.ndf
    assert_bits(cond.dict_Step["SYSVAR"][:, 0], ref["sysvar"], "SYSVAR")
    assert_bits(cond.dict_Step["SYSLOD"], ref["syslod"], "SYSLOD")
    assert_bits(cond.dict_norm["Solution"], ref["norm_solution"], "norm")
    assert_bits(cond.dict_norm["Change"], ref["norm_change"], "change")
    assert_bits(cond.EQTEIG, ref["eqteig"], "EQTEIG")
    for j, f in enumerate(cond.inventory["FluidComponent"].collection):
        assert_bits(f.coolant.dict_node_pt["velocity"], ref["sysvar"][j::d], "v")
        assert_bits(f.coolant.dict_node_pt["pressure"], ref["sysvar"][M + j::d], "p")
        t = ref["sysvar"][2 * M + j::d]
        assert_bits(f.coolant.dict_node_pt["temperature"], t, "T")
        assert_bits(f.coolant.dict_Gauss_pt["temperature_change"],
                    (t[:-1] + t[1:]) / 2. - arr.fl_gauss[2, j], "dT")
    for l, s in enumerate(cond.inventory["SolidComponent"].collection):
        assert_bits(s.dict_node_pt["temperature"], ref["sysvar"][3 * M + l::d], "Ts")
    for k, i in enumerate(cond.interface.fluid_fluid):
        assert_bits(cond.dict_Gauss_pt["K1"][i.interf_name], ref["k123"][0, k], "K1")
        assert_bits(cond.dict_Gauss_pt["K3"][i.interf_name], ref["k123"][2, k], "K3")


def test_unsupported_method_is_rejected():
    topo, arr, _ = load_golden("case1_pdrop_n21")
    cond, env, qsource = build_conductor(topo, arr)
    cond.inputs["METHOD"] = "AM4"
    with pytest.raises(ValueError):
        shim.topology_of(cond)


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the shim must raise, never compute on the host."""
    import ctypes as C

    from opensc2_b200 import _lib

    L = _lib.load()
    if L.osc2_device_count() > 0:
        pytest.skip("a CUDA device is present")
    topo, arr, _ = load_golden("case1_pdrop_n21")
    cond, env, qsource = build_conductor(topo, arr)
    before = cond.dict_Step["SYSVAR"].copy()
    with pytest.raises(_lib.Osc2Error) as e:
        shim.step(cond, env, qsource, arr.num_step)
    assert e.value.code == -3
    assert np.array_equal(cond.dict_Step["SYSVAR"], before)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["case1_pdrop_n21", "case1_pinl_vout_first_step_n17",
                                  "case1_vinl_pout_cn_refined_n12", "case3_backward_n15",
                                  "case3_rectangular_env_n9", "case2_n6"])
def test_step_shim_bit_exact_vs_reference_fixture(name):
    topo, arr, ref = load_golden(name)
    cond, env, qsource = build_conductor(topo, arr)
    assert shim.step(cond, env, qsource, arr.num_step) is None
    d, M = topo.ndf, topo.n_ch
    assert_bits(cond.dict_Step["SYSVAR"][:, 0], ref["sysvar"], "SYSVAR")
    assert_bits(cond.dict_Step["SYSLOD"], ref["syslod"], "SYSLOD")
    assert_bits(cond.dict_norm["Solution"], ref["norm_solution"], "norm")
    assert_bits(cond.dict_norm["Change"], ref["norm_change"], "change")
    assert_bits(cond.EQTEIG, ref["eqteig"], "EQTEIG")
    for j, f in enumerate(cond.inventory["FluidComponent"].collection):
        assert_bits(f.coolant.dict_node_pt["pressure"], ref["sysvar"][M + j::d], "p")
    shim.release()


@pytest.mark.gpu
def test_step_many_equals_one_by_one():
    from opensc2_b200.problem import case1_topology, synthetic_step_arrays, take

    topo = case1_topology(3)
    arr = synthetic_step_arrays(topo, 33, batch=5, seed=9)
    many = [build_conductor(topo, take(arr, b)) for b in range(5)]
    single = [build_conductor(topo, take(arr, b)) for b in range(5)]
    shim.step_many([m[0] for m in many], many[0][1], [m[2] for m in many], arr.num_step)
    for (c1, env, qs), (cm, _e, _q) in zip(single, many):
        shim.step(c1, env, qs, arr.num_step)
        assert_bits(cm.dict_Step["SYSVAR"], c1.dict_Step["SYSVAR"], "SYSVAR")
        assert_bits(cm.dict_norm["Change"], c1.dict_norm["Change"], "Change")
    shim.release()
