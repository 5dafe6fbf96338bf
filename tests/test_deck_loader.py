"""Deck loader (next-1): the shipped TDD_examples workbooks -> Topology / Model.

* With the decks present (build container): `load_deck` must reproduce the committed summary
  (tests/golden/decks_summary.json, written by make_golden_decks.py).
* Always: the hand-written CASE_1 / CASE_3 topologies and models used by the benchmarks and parity tests
  equal what the decks say (interfaces, areas, materials, HTC tables), and the counts of SURVEY.md
  Appendix A hold (d = 8 / 64 / 13; 5 / 97 / 9 interfaces)."""
import dataclasses
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN_DIR
from opensc2_b200.model import case1_model
from opensc2_b200.problem import case1_topology, case3_topology

DECKS = "/root/reference/TDD_examples"
SUMMARY = json.load(open(os.path.join(GOLDEN_DIR, "decks_summary.json")))


def _norm(x):
    if isinstance(x, dict):
        return {k: _norm(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_norm(v) for v in x]
    return x


@pytest.mark.skipif(not os.path.isdir(DECKS), reason="the reference decks are not on this machine")
@pytest.mark.parametrize("case,override", [("CASE_1_ITER_like_LTS", 1), ("CASE_2_ENEA_HTS_CICC", 3), ("CASE_3_HTS_HVDC", 3)])
def test_loader_reproduces_committed_summary(case, override):
    from opensc2_b200.deck import load_deck

    d = load_deck(os.path.join(DECKS, case), intial_override=override)
    got = _norm({"topology": dataclasses.asdict(d.topology), "model": dataclasses.asdict(d.model),
                 "chan_ids": d.chan_ids, "solid_ids": d.solid_ids, "length": d.length, "n_elems": d.n_elems,
                 "fl_bc": d.fl_bc.tolist(), "peri": {k: v.tolist() for k, v in d.peri.items()}})
    want = SUMMARY[case]
    for key, v in got.items():
        assert json.loads(json.dumps(v)) == want[key], key
    d.topology.validate()
    d.model.struct(d.topology)                     # fits the C descriptor


def test_counts_of_the_three_decks():
    sizes = {c: (v["topology"]["n_ch"] * 3 + v["topology"]["n_sol"],
                 sum(len(v["topology"][k]) for k in ("ff", "fs", "ss", "es"))) for c, v in SUMMARY.items()}
    assert sizes == {"CASE_1_ITER_like_LTS": (8, 5), "CASE_2_ENEA_HTS_CICC": (64, 97), "CASE_3_HTS_HVDC": (13, 9)}


def test_hand_written_case1_matches_the_deck():
    t, want = dataclasses.asdict(case1_topology(1)), SUMMARY["CASE_1_ITER_like_LTS"]["topology"]
    for key in ("n_ch", "n_sol", "ch_area", "ch_dh", "ch_forward", "sol_area", "sol_costeta", "ff", "fs", "ss", "es"):
        assert json.loads(json.dumps(t[key])) == want[key], key
    m, wm = dataclasses.asdict(case1_model(False)), SUMMARY["CASE_1_ITER_like_LTS"]["model"]
    for key in ("ch_ifriction", "ch_friction_mult", "ch_htc_corr", "solids", "fs", "ss", "es_htc"):
        assert json.loads(json.dumps(m[key])) == wm[key], key
    assert [c[:3] for c in m["ff"]] == [tuple(c[:3]) for c in wm["ff"]] or [list(c[:3]) for c in m["ff"]] == [c[:3] for c in wm["ff"]]


def test_hand_written_case3_interfaces_match_the_deck():
    t, want = dataclasses.asdict(case3_topology()), SUMMARY["CASE_3_HTS_HVDC"]["topology"]
    for key in ("n_ch", "n_sol", "ch_forward", "ff", "fs", "ss", "es"):
        assert json.loads(json.dumps(t[key])) == want[key], key
    for key in ("ch_area", "ch_dh"):               # typed from the survey's rounded table
        assert np.allclose(t[key], want[key], rtol=1e-3), key
    choices = [c[0] for c in SUMMARY["CASE_3_HTS_HVDC"]["model"]["ss"]]
    assert choices == [3, -1, 3, -1]               # two radiative gaps, two contacts


def test_head_schema_stack_column_becomes_a_stack_solid():
    """STACK sheet column of the HEAD schema -> SOLID_STACK with the layers in sheet order ("none" materials
    and zero thicknesses dropped like StackComponent.__reorganize_input does)."""
    from opensc2_b200 import model as M
    from opensc2_b200.deck import stack_model

    col = {"CROSSECTION": 4.8e-6, "RRR": 80.0, "Stack_width": 4e-3, "Tape_number": 12,
           "HTS_material": "RE123", "HTS_thickness": 1.0e-6,
           "Stabilizer_material": "Cu", "Stabilizer_thickness": 40.0e-6,
           "Substrate_material": "HC276", "Substrate_thickness": 50.0e-6,
           "Silver_material": "Ag", "Silver_thickness": 2.0e-6,
           "Solder_material": "none", "Solder_thickness": 0.0}
    sm = stack_model(col, heated=True)
    assert sm.kind == M.SOLID_STACK and sm.heated and sm.rrr == 80.0
    assert sm.materials == (M.MAT_RE123, M.MAT_CU_NIST, M.MAT_HC276, M.MAT_AG)
    assert sm.weights == (1.0e-6, 40.0e-6, 50.0e-6, 2.0e-6)
    assert sm.weight_sum == float(np.array(sm.weights).sum())
    assert stack_model({"CROSSECTION": 1.0}) is None
    with pytest.raises(ValueError):
        stack_model({"HTS_material": "RE123", "HTS_thickness": 0.0})
