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
    from opensc2_b200.deck import deck_from_dict, deck_to_dict

    d = load_deck(os.path.join(DECKS, case), intial_override=override)
    got, want = deck_to_dict(d), SUMMARY[case]
    for key, v in got.items():
        assert json.loads(json.dumps(v)) == want[key], key
    e = deck_from_dict(want)                         # the committed image rebuilds the same deck (GPU box: no workbooks)
    assert e.topology == d.topology and e.model == d.model and np.array_equal(e.fl_bc, d.fl_bc)
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


# ---------------------------------------------------------------------------------------------------
# Independent pins (VERDICT r1 "weak" 3): values read from the workbooks' XML by a reader written HERE
# (regular expressions over the sheet XML, nothing shared with opensc2_b200.xlsx / deck.py), and the three
# legacy inferences checked against what the reference's own golden `Solution/*.tsv` files say.
# ---------------------------------------------------------------------------------------------------
def _raw_cells(path, sheet_name):
    """{cell ref: value} of one sheet; strings resolved through sharedStrings."""
    import re
    import zipfile

    with zipfile.ZipFile(path) as z:
        shared = [re.sub(r"<[^>]+>", "", si) for si in
                  re.findall(r"<si>(.*?)</si>", z.read("xl/sharedStrings.xml").decode("utf8"), flags=re.S)]
        wb = z.read("xl/workbook.xml").decode("utf8")
        rels = dict(re.findall(r'<Relationship[^>]*?Id="([^"]+)"[^>]*?Target="([^"]+)"', z.read("xl/_rels/workbook.xml.rels").decode("utf8")))
        rels.update({i: t for t, i in re.findall(r'<Relationship[^>]*?Target="([^"]+)"[^>]*?Id="([^"]+)"', z.read("xl/_rels/workbook.xml.rels").decode("utf8"))})
        rid = re.search(r'<sheet[^>]*name="%s"[^>]*r:id="([^"]+)"' % re.escape(sheet_name), wb).group(1)
        target = rels[rid].lstrip("/")
        xml = z.read(target if target.startswith("xl/") else "xl/" + target).decode("utf8")
    cells = {}
    for attrs, body in re.findall(r"<c ([^>]*?)(?:/>|>(.*?)</c>)", xml, flags=re.S):
        ref = re.search(r'r="([A-Z]+[0-9]+)"', attrs).group(1)
        v = re.search(r"<v>(.*?)</v>", body or "")
        if v is None:
            continue
        t = re.search(r't="(\w+)"', attrs)
        cells[ref] = shared[int(v.group(1))] if t and t.group(1) == "s" else (v.group(1) if t and t.group(1) == "str" else float(v.group(1)))
    return cells


def _raw_variable(path, sheet, variable, column_header):
    """Value in the row whose first text cell is `variable`, in the column whose header row holds `column_header`."""
    import re

    cells = _raw_cells(path, sheet)
    col = [re.match(r"[A-Z]+", r).group(0) for r, v in cells.items() if v == column_header][0]
    row = [re.search(r"[0-9]+", r).group(0) for r, v in cells.items() if v == variable and r.startswith("A")][0]
    return cells[col + row]


@pytest.mark.skipif(not os.path.isdir(DECKS), reason="the reference decks are not on this machine")
def test_loader_against_an_independent_reading_of_the_workbooks():
    from opensc2_b200.deck import load_deck

    c1 = os.path.join(DECKS, "CASE_1_ITER_like_LTS")
    d = load_deck(c1, intial_override=1)
    inp, op = os.path.join(c1, "conductor_1_input.xlsx"), os.path.join(c1, "conductor_1_operation.xlsx")
    for j, ch in enumerate(("CHAN_1", "CHAN_2")):
        assert d.topology.ch_area[j] == _raw_variable(inp, "CHAN", "CROSSECTION", ch)
        assert d.topology.ch_dh[j] == _raw_variable(inp, "CHAN", "HYDIAMETER", ch)
        assert d.model.ch_ifriction[j] == int(_raw_variable(inp, "CHAN", "IFRICTION", ch))
        assert d.model.ch_friction_mult[j] == _raw_variable(inp, "CHAN", "FRICTION_MULTIPLIER", ch)
        for row, key in enumerate(("PREINL", "PREOUT", "TEMINL", "TEMOUT", "MDTIN")):
            assert d.fl_bc[row, j] == _raw_variable(op, "CHAN", key, ch), (ch, key)
    sm = d.model.solids[0]
    assert sm.weights[0] == _raw_variable(inp, "STR_MIX", "STAB_NON_STAB", "STR_MIX_1")
    assert (sm.rrr, sm.tc0m, sm.bc20m, sm.c0) == tuple(_raw_variable(inp, "STR_MIX", k, "STR_MIX_1") for k in ("RRR", "Tc0m", "Bc20m", "c0"))
    assert d.topology.sol_area[0] == _raw_variable(inp, "STR_MIX", "CROSSECTION", "STR_MIX_1")
    assert d.topology.sol_costeta[0] == _raw_variable(inp, "STR_MIX", "COSTETA", "STR_MIX_1")
    jk = d.model.solids[1]
    assert jk.weights == (_raw_variable(inp, "Z_JACKET", "CROSSECTION_JK", "Z_JACKET_1"), _raw_variable(inp, "Z_JACKET", "CROSSECTION_IN", "Z_JACKET_1"))
    assert d.solid_ops[0]["Q0"] == _raw_variable(op, "STR_MIX", "Q0", "STR_MIX_1")
    for key in ("XQBEG", "XQEND", "TQBEG", "TQEND"):
        assert d.solid_ops[0][key] == _raw_variable(op, "STR_MIX", key, "STR_MIX_1"), key
    assert d.n_elems == int(_raw_variable(os.path.join(c1, "conductor_grid.xlsx"), "GRID", "NELEMS", "CONDUCTOR_1"))
    assert d.length == _raw_variable(os.path.join(c1, "conductor_definition.xlsx"), "CONDUCTOR_input", "XLENGTH", "CONDUCTOR_1")
    # CASE_3: second fluid, flow direction, a radiative pair
    c3 = os.path.join(DECKS, "CASE_3_HTS_HVDC")
    d3 = load_deck(c3, intial_override=3)
    inp3, op3 = os.path.join(c3, "conductor_1_input.xlsx"), os.path.join(c3, "conductor_1_operation.xlsx")
    assert str(_raw_variable(inp3, "CHAN", "FLUID_TYPE", "CHAN_2")).lower().startswith("n") and d3.model.ch_fluid == (0, 1)
    assert str(_raw_variable(op3, "CHAN", "FLOWDIR", "CHAN_2")).lower() == "backward" and d3.topology.ch_forward == (True, False)
    assert d3.fl_bc[4, 1] == _raw_variable(op3, "CHAN", "MDTIN", "CHAN_2")


@pytest.mark.skipif(not os.path.isdir(DECKS), reason="the reference decks are not on this machine")
def test_the_three_legacy_inferences_against_the_golden_solutions():
    """(i) hydraulic boundary mode: CASE_1's golden keeps p(0) and p(L) at the deck's PREINL / PREOUT and lets the
    mass flow float (-> INTIAL 1); CASE_2 / CASE_3 keep the inlet mass flow and temperature (-> INTIAL 3, outlet
    pressure = the golden's).  (ii) legacy IBIFUN = 0 meant no field: the golden B_field column is zero although
    BISS = BOSS = 1 in the deck.  (iii) legacy STR_SC files carry the superconducting-strand layout (4 columns)."""
    from opensc2_b200.deck import load_deck

    def tsv(case, name):
        return np.loadtxt(os.path.join(DECKS, case, "Solution", name + ".tsv"), skiprows=1)

    d1 = load_deck(os.path.join(DECKS, "CASE_1_ITER_like_LTS"), intial_override=1)
    for j, ch in enumerate(d1.chan_ids):
        g = tsv("CASE_1_ITER_like_LTS", ch)
        assert g[0, 1] == d1.fl_bc[0, j] and g[-1, 1] == d1.fl_bc[1, j] and g[0, 2] == d1.fl_bc[2, j]
        assert abs(g[0, -1] - d1.fl_bc[4, j]) > 1e-7                     # the mass flow is NOT the deck's: it floats
    g = tsv("CASE_1_ITER_like_LTS", "STR_MIX_1")
    assert np.all(g[:, 2] == 0.0)                                         # B_field
    assert any("no field" in n for n in d1.notes) and d1.solid_ops[0]["BISS"] == 0.0
    for case in ("CASE_2_ENEA_HTS_CICC", "CASE_3_HTS_HVDC"):
        d = load_deck(os.path.join(DECKS, case), intial_override=3)
        for j, ch in enumerate(d.chan_ids):
            g = tsv(case, ch)
            inlet = 0 if d.topology.ch_forward[j] else -1
            assert abs(abs(g[inlet, -1]) - d.fl_bc[4, j]) <= 1e-7 * d.fl_bc[4, j], (case, ch)   # inlet mass flow kept (density lagged one step)
            assert g[inlet, 2] == d.fl_bc[2, j]                                                        # inlet temperature kept
    assert tsv("CASE_2_ENEA_HTS_CICC", "STR_SC_1").shape[1] == 4
