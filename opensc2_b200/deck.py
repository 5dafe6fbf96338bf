"""Deck loader: turns a folder of OPENSC2 input workbooks into the flat
`Topology` / `Model` / operating data the batched device path consumes.

The reference does this in `Conductor.__init__` and its helpers (conductor.py:99-306, 680-863,
1354-1830, 2554-2675) with openpyxl + pandas; here the workbooks are read with the standard library
(`xlsx.py`).  Both the legacy schema of the shipped `TDD_examples` decks and the key names of HEAD are
accepted (SURVEY.md Appendix B lists the renames); the three inferences the legacy decks need are
explicit arguments, not silent choices:
  * `legacy_no_field`: in the legacy decks `IBIFUN == 0` meant "no field" (the CASE_1 golden has
    B = 0 although BISS = BOSS = 1), HEAD would take the linspace;
  * `intial_override`: HEAD's hydraulic boundary mode (1, 2, 3) that reproduces the effective
    boundary conditions of the goldens (CASE_1 -> 1, CASE_2/3 -> 3);
  * legacy `STR_SC` sheets (CASE_2) become single-material RE123 solids.
Nothing here runs per time step.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import xlsx
from .model import (MAT_AG, MAT_AL, MAT_CU_NIST, MAT_GE, MAT_HC276, MAT_MGB2, MAT_NB3SN, MAT_NBTI, MAT_RE123, MAT_SN60PB40, MAT_SS,
                    SOLID_STACK, SOLID_JACKET, SOLID_STRAND_MIXED,
                    SOLID_STRAND_STAB, Model, SolidModel, radiative_pair)
from .problem import Topology

_MATERIALS = {"cu": MAT_CU_NIST, "copper": MAT_CU_NIST, "nb3sn": MAT_NB3SN, "ss": MAT_SS, "steinless_steel": MAT_SS,
              "stainless_steel": MAT_SS, "ge": MAT_GE, "glass_epoxy": MAT_GE, "al": MAT_AL, "aluminium": MAT_AL,
              "hts": MAT_RE123, "re123": MAT_RE123, "ybco": MAT_RE123, "nbti": MAT_NBTI, "mgb2": MAT_MGB2,
              "ag": MAT_AG, "silver": MAT_AG, "hc276": MAT_HC276, "sn60pb40": MAT_SN60PB40}


@dataclass
class Deck:
    folder: str
    topology: Topology
    model: Model
    chan_ids: List[str]
    solid_ids: List[str]
    length: float
    n_elems: int
    mesh: Dict[str, float]
    transient: Dict[str, object]
    t_env: float
    fl_bc: np.ndarray                      # (6, M) PREINL, PREOUT, TEMINL, TEMOUT, MDTIN, MDTOUT
    solid_ops: List[Dict[str, object]]     # per solid: BISS, BOSS, EPS, IQFUN, Q0, XQBEG, XQEND, TQBEG, TQEND, INTIAL, TEMINL, TEMOUT
    peri: Dict[str, np.ndarray]            # constant contact perimeters per interface list: ff (2, nff), fs, ss, es
    notes: List[str] = field(default_factory=list)
    solid_inputs: List[Dict[str, object]] = field(default_factory=list)   # raw input-sheet column of every solid
    ss_view_factor: List[float] = field(default_factory=list)             # per solid-solid interface


def deck_to_dict(d: "Deck") -> dict:
    """Plain-data image of a Deck (JSON-able): lets a parsed deck travel without its workbooks."""
    import dataclasses

    def clean(x):
        if isinstance(x, dict):
            return {k: clean(v) for k, v in x.items()}
        if isinstance(x, (list, tuple)):
            return [clean(v) for v in x]
        if isinstance(x, np.ndarray):
            return x.tolist()
        if isinstance(x, float) and x == float("inf"):
            return "inf"
        if isinstance(x, (np.floating, np.integer)):
            return x.item()
        return x

    return clean({"topology": dataclasses.asdict(d.topology), "model": dataclasses.asdict(d.model), "chan_ids": d.chan_ids,
                  "solid_ids": d.solid_ids, "length": d.length, "n_elems": d.n_elems, "mesh": d.mesh, "t_env": d.t_env,
                  "fl_bc": d.fl_bc, "solid_ops": d.solid_ops, "peri": d.peri,
                  "transient": {k: v for k, v in d.transient.items() if isinstance(v, (int, float, str))}, "notes": d.notes,
                  "solid_inputs": d.solid_inputs, "ss_view_factor": d.ss_view_factor})


def deck_from_dict(x: dict, folder: str = "") -> "Deck":
    def tup(v):
        return tuple(tup(i) for i in v) if isinstance(v, list) else (float("inf") if v == "inf" else v)

    t = {k: tup(v) for k, v in x["topology"].items()}
    m = dict(x["model"])
    solids = tuple(SolidModel(**{k: tup(v) for k, v in s.items()}) for s in m.pop("solids"))
    model = Model(solids=solids, **{k: tup(v) for k, v in m.items()})
    peri = {k: np.array(v, dtype=np.float64).reshape((2, -1) if k == "ff" else (-1,)) for k, v in x["peri"].items()}
    return Deck(folder=folder, topology=Topology(**t), model=model, chan_ids=list(x["chan_ids"]), solid_ids=list(x["solid_ids"]),
                length=float(x["length"]), n_elems=int(x["n_elems"]), mesh=dict(x.get("mesh", {})), transient=dict(x["transient"]),
                t_env=float(x["t_env"]), fl_bc=np.array(x["fl_bc"], dtype=np.float64), solid_ops=list(x["solid_ops"]), peri=peri,
                notes=list(x.get("notes", [])), solid_inputs=list(x.get("solid_inputs", [])),
                ss_view_factor=list(x.get("ss_view_factor", [])))


def _objects(sheet_rows, kind):
    tab = xlsx.variable_table(sheet_rows)
    n = int(sheet_rows[0][1]) if sheet_rows and len(sheet_rows[0]) > 1 and isinstance(sheet_rows[0][1], float) else 0
    ids = [k for k in tab if k.startswith(kind + "_") and k.split("_")[-1].isdigit() and int(k.split("_")[-1]) >= 1]
    ids.sort(key=lambda s: int(s.split("_")[-1]))
    return {i: tab[i] for i in ids[:n]}


def _get(d, *names, default=None):
    for n in names:
        if n in d and d[n] is not None:
            return d[n]
    return default


def stack_model(d, heated: bool = False):
    """SolidModel of a HEAD-schema STACK sheet column (StackComponent.__reorganize_input, stack_component.py:216-
    272): the `*material` entries that are not "none" are the tape layers, in sheet order; the non-zero
    `*thickness` entries are their thicknesses; the stack is homogenised by thickness.  None when the column
    carries no `*material` keys (then it is not a HEAD stack)."""
    mats = [str(v).lower() for k, v in d.items() if str(k).endswith("material") and str(v).lower() != "none"]
    thick = [float(v) for k, v in d.items() if str(k).endswith("thickness") and float(v) != 0.0]
    if not mats:
        return None
    if len(mats) != len(thick):
        raise ValueError(f"STACK: {len(mats)} tape materials but {len(thick)} non-zero layer thicknesses "
                         "(the reference raises in StackComponent.__check_consistency)")
    return SolidModel(SOLID_STACK, tuple(_material(m) for m in mats), tuple(thick), float(np.array(thick).sum()),
                      rrr=float(d.get("RRR", 100.0)), heated=heated)


def _material(name) -> int:
    key = str(name).strip().lower()
    if key not in _MATERIALS:
        raise ValueError(f"material {name!r} is not implemented on the device")
    return _MATERIALS[key]


def load_deck(folder: str, conductor: int = 1, legacy_no_field: bool = True,
              intial_override: Optional[int] = None) -> Deck:
    notes: List[str] = []
    definition = xlsx.read_workbook(os.path.join(folder, "conductor_definition.xlsx"))
    files = xlsx.variable_table(definition["CONDUCTOR_files"])[f"CONDUCTOR_{conductor}"]
    cin = xlsx.variable_table(definition["CONDUCTOR_input"])[f"CONDUCTOR_{conductor}"]
    inputs = xlsx.read_workbook(os.path.join(folder, files["STRUCTURE_ELEMENTS"]))
    oper = xlsx.read_workbook(os.path.join(folder, files["OPERATION"]))
    coup = {k: xlsx.matrix_table(v) for k, v in xlsx.read_workbook(os.path.join(folder, files["STRUCTURE_COUPLING"])).items()}
    grid = xlsx.variable_table(xlsx.read_workbook(os.path.join(folder, files["GRID_DEFINITION"]))["GRID"])[f"CONDUCTOR_{conductor}"]
    transient_rows = xlsx.read_workbook(os.path.join(folder, "transitory_input.xlsx"))["TRANSIENT"]
    transient = {r[0]: r[4] for r in transient_rows if len(r) > 4 and isinstance(r[0], str) and r[0] != "Variable name"}
    env_rows = xlsx.read_workbook(os.path.join(folder, str(transient.get("ENVIRONMENT", "environment_input.xlsx"))))["ENVIRONMENT"]
    env = {r[0]: r[4] for r in env_rows if len(r) > 4 and isinstance(r[0], str)}

    # ---- components, in the reference's inventory order: channels, then strands (STR_MIX, STR_SC / STACK,
    # STR_STAB), then jackets (conductor.py:680-863)
    chans = _objects(inputs["CHAN"], "CHAN")
    chan_ops = _objects(oper["CHAN"], "CHAN")
    solids: List[Tuple[str, str, dict, dict]] = []
    for kind in ("STR_MIX", "STR_SC", "STACK", "STR_STAB", "Z_JACKET"):
        if kind in inputs:
            objs, ops = _objects(inputs[kind], kind), _objects(oper[kind], kind) if kind in oper else {}
            for ident, d in objs.items():
                solids.append((kind, ident, d, ops.get(ident, {})))
    chan_ids, solid_ids = list(chans), [s[1] for s in solids]
    M, S = len(chan_ids), len(solid_ids)

    ch_area = tuple(float(chans[c]["CROSSECTION"]) for c in chan_ids)
    ch_dh = tuple(float(chans[c]["HYDIAMETER"]) for c in chan_ids)
    ch_forward = tuple(str(_get(chan_ops[c], "FLOWDIR", default="forward")).lower() == "forward" for c in chan_ids)
    deck_intial = [int(chan_ops[c]["INTIAL"]) for c in chan_ids]
    if any(i not in (1, 2, 3) for i in deck_intial):
        raise ValueError(f"INTIAL {deck_intial}: only 1, 2, 3 are implemented (negative flags read time-dependent "
                         "boundary values from the EXTERNAL_FLOW workbook)")
    if intial_override is not None:
        notes.append(f"hydraulic boundary mode INTIAL {sorted(set(deck_intial))} of the deck replaced by {intial_override}")
    ch_intial = tuple(intial_override if intial_override is not None else i for i in deck_intial)
    fl_bc = np.zeros((6, M))
    for j, c in enumerate(chan_ids):
        o = chan_ops[c]
        mdt_in = float(_get(o, "MDTIN", default=0.0))
        fl_bc[:, j] = (float(o["PREINL"]), float(o["PREOUT"]), float(o["TEMINL"]), float(o["TEMOUT"]), mdt_in,
                       float(_get(o, "MDTOUT", default=mdt_in)))

    solid_models, sol_area, sol_cos, solid_ops = [], [], [], []
    for kind, ident, d, o in solids:
        cos = float(_get(d, "COSTETA", default=1.0))
        heated = int(_get(o, "IQFUN", default=0)) == 1
        if kind == "STR_MIX":
            stab = _material(_get(d, "stabilizer_material", "ISTABILIZER"))
            sc = _material(_get(d, "superconducting_material", "ISUPERCONDUCTOR"))
            ratio = float(d["STAB_NON_STAB"])
            sm = SolidModel(SOLID_STRAND_MIXED, (stab, sc), (ratio, 1.0), float(np.array([ratio, 1.0]).sum()),
                            rrr=float(_get(d, "RRR", default=100.0)), tc0m=float(_get(d, "Tc0m", default=0.0)),
                            bc20m=float(_get(d, "Bc20m", default=0.0)), heated=heated,
                            c0=float(_get(d, "c0", "C0", default=0.0)))
            area = float(d["CROSSECTION"])
        elif kind in ("STR_SC", "STACK"):
            sm = stack_model(d, heated) if kind == "STACK" else None
            if sm is None:
                if kind == "STR_SC":
                    notes.append(f"{ident}: legacy STR_SC restated as a single-material RE123 solid (SURVEY.md Appendix B)")
                sm = SolidModel(SOLID_STRAND_STAB, (MAT_RE123,), (1.0,), 1.0, heated=heated)
            area = float(d["CROSSECTION"])
        elif kind == "STR_STAB":
            mat = _material(_get(d, "stabilizer_material", "ISTABILIZER"))
            sm = SolidModel(SOLID_STRAND_STAB, (mat,), (1.0,), 1.0, rrr=float(_get(d, "RRR", default=100.0)), heated=heated)
            area = float(d["CROSSECTION"])
        else:
            a_jk = float(_get(d, "jacket_cross_section", "CROSSECTION_JK"))
            a_in = float(_get(d, "insulation_cross_section", "CROSSECTION_IN", default=0.0))
            m_jk = _get(d, "jacket_material", "IMATERIAL_JK")
            m_in = _get(d, "insulation_material", "IMATERIAL_IN", default="none")
            mats, ws = [], []                                   # jacket_component.py:231-262: "none" / zero area dropped
            for mname, a in ((m_jk, a_jk), (m_in, a_in)):
                if str(mname).strip().lower() != "none" and a != 0.0:
                    mats.append(_material(mname)); ws.append(a)
            area = a_jk + a_in                                   # jacket_component.py:300-303
            sm = SolidModel(SOLID_JACKET, tuple(mats), tuple(ws), area, heated=heated)
        solid_models.append(sm); sol_area.append(area); sol_cos.append(cos)
        biss, boss = float(_get(o, "BISS", default=0.0)), float(_get(o, "BOSS", default=0.0))
        legacy_schema = "stabilizer_material" not in d and "jacket_material" not in d
        if legacy_no_field and legacy_schema and int(_get(o, "IBIFUN", default=0)) == 0 and (biss or boss):
            notes.append(f"{ident}: legacy IBIFUN = 0 read as 'no field' (BISS = {biss}, BOSS = {boss} ignored)")
            biss = boss = 0.0
        solid_ops.append(dict(BISS=biss, BOSS=boss, EPS=float(_get(o, "EPS", default=0.0)) if int(_get(o, "IEPS", default=0)) == 1 else 0.0,
                              IQFUN=int(_get(o, "IQFUN", default=0)), Q0=float(_get(o, "Q0", default=0.0)),
                              XQBEG=float(_get(o, "XQBEG", default=0.0)), XQEND=float(_get(o, "XQEND", default=0.0)),
                              TQBEG=float(_get(o, "TQBEG", default=0.0)), TQEND=float(_get(o, "TQEND", default=0.0)),
                              INTIAL=int(_get(o, "INTIAL", default=0)), TEMINL=float(_get(o, "TEMINL", default=0.0)),
                              TEMOUT=float(_get(o, "TEMOUT", default=0.0))))

    # ---- interfaces from the upper triangle of contact_perimeter_flag, in the reference's loop order
    flag, perim, choice = coup["contact_perimeter_flag"], coup["contact_perimeter"], coup["HTC_choice"]
    htc, mlt = coup["contact_HTC"], coup["HTC_multiplier"]
    thick = coup.get("interf_thickness", {})
    rcont = coup.get("thermal_contact_resistance", {})
    openf = coup.get("open_perimeter_fract", {})
    view = coup.get("view_factors", {})
    at = lambda tab, a, b, default=0.0: float(tab.get(a, {}).get(b, default) or 0.0)
    ff, fs, ss, es = [], [], [], []
    m_fs, m_ff, m_ss, m_ss_rad, m_es = [], [], [], [], []
    p_ff, p_fs, p_ss, p_es = [], [], [], []
    v_ss: List[float] = []
    phi_conv = float(_get(cin, "Phi_conv", default=1.0))
    for j, a in enumerate(chan_ids):
        for l, b in enumerate(solid_ids):
            if abs(at(flag, a, b)) == 1:
                fs.append((j, l)); m_fs.append((int(at(choice, a, b)), at(htc, a, b), at(mlt, a, b, 1.0))); p_fs.append(at(perim, a, b))
        for j2 in range(j + 1, M):
            b = chan_ids[j2]
            if abs(at(flag, a, b)) == 1:
                ff.append((j, j2)); m_ff.append((int(at(choice, a, b)), at(htc, a, b), at(mlt, a, b, 1.0), at(thick, a, b)))
                frac = at(openf, a, b)
                p_ff.append((at(perim, a, b) * frac, at(perim, a, b) * (1.0 - frac)))     # conductor.py:1655-1677
    for l, a in enumerate(solid_ids):
        for l2 in range(l + 1, S):
            b = solid_ids[l2]
            if abs(at(flag, a, b)) == 1:
                ch = int(at(choice, a, b))
                ss.append((l, l2)); p_ss.append(at(perim, a, b)); v_ss.append(at(view, a, b))
                m_ss.append((ch, at(htc, a, b), at(mlt, a, b, 1.0), at(thick, a, b), at(thick, b, a), at(rcont, a, b)))
                if ch == 3:
                    da, db = solids[l][2], solids[l2][2]
                    m_ss_rad.append(radiative_pair(float(_get(da, "Emissivity", default=0.0)), float(_get(db, "Emissivity", default=0.0)),
                                                   float(_get(da, "Outer_perimeter", default=0.0)), float(_get(da, "Inner_perimeter", default=0.0)),
                                                   float(_get(db, "Outer_perimeter", default=0.0)), float(_get(db, "Inner_perimeter", default=0.0)),
                                                   at(view, a, b), at(mlt, a, b, 1.0)))
                else:
                    m_ss_rad.append((float("inf"), 1.0))
        if abs(at(flag, "Environment", a)) == 1:
            ch = int(at(choice, "Environment", a))
            if ch not in (-2, -3, 3, -4):
                raise ValueError(f"Environment-{a}: HTC_choice {ch} (free convection from air properties) is not "
                                 "implemented on the device (-2, 3, -3, -4 are)")
            es.append(l); p_es.append(at(perim, "Environment", a))
            if abs(ch) == 3:
                # radiation only: the reference stores the coefficient under "rad" (conductor.py:5320-5350), and the
                # step reads "conv" alone (smc:851-853, 969-971), which stays zero -- no exchange reaches the system
                h = 0.0
                notes.append(f"Environment-{a}: HTC_choice {ch}: radiative coefficient is output-only in the reference; "
                             "the step sees a zero convective coefficient")
            else:
                h = at(htc, "Environment", a) * (phi_conv if ch == -2 else 1.0) * at(mlt, "Environment", a, 1.0)   # conductor.py:5302-5386
            m_es.append(h)

    method = str(_get(cin, "METHOD", default="BE"))
    topo = Topology(n_ch=M, n_sol=S, ch_area=ch_area, ch_dh=ch_dh, ch_forward=ch_forward, ch_intial=ch_intial,
                    sol_area=tuple(sol_area), sol_costeta=tuple(sol_cos), ff=tuple(ff), fs=tuple(fs), ss=tuple(ss),
                    es=tuple(es), es_choice2=(False,) * len(es), is_rectangular=bool(_get(cin, "Is_rectangular", default=False)),
                    height=float(_get(cin, "Height", default=0.0)), width=float(_get(cin, "Width", default=0.0)),
                    theta={"BE": 1.0, "CN": 0.5}.get(method, 1.0), method=method)
    model = Model(ch_fluid=tuple(0 if str(_get(chans[c], "FLUID_TYPE", default="Helium")).lower().startswith("he") else 1 for c in chan_ids),
                  ch_ifriction=tuple(int(chans[c]["IFRICTION"]) for c in chan_ids),
                  ch_friction_mult=tuple(float(_get(chans[c], "FRICTION_MULTIPLIER", default=1.0)) for c in chan_ids),
                  ch_htc_corr=tuple(int(_get(chans[c], "Flag_htc_steady_corr", default=1)) for c in chan_ids),
                  ch_roughness=tuple(float(_get(chans[c], "Roughness", default=0.0)) for c in chan_ids),
                  ch_void_fraction=tuple(float(_get(chans[c], "VOID_FRACTION", default=0.0)) for c in chan_ids),
                  ch_side_ratio=tuple((float(_get(chans[c], "SIDE2", default=0.0)) / float(_get(chans[c], "SIDE1", default=1.0)))
                                      if float(_get(chans[c], "SIDE1", default=0.0)) != 0.0 else 0.0 for c in chan_ids),
                  solids=tuple(solid_models), fs=tuple(m_fs), ff=tuple(m_ff), ss=tuple(m_ss), ss_rad=tuple(m_ss_rad),
                  es_htc=tuple(m_es))
    peri = {"ff": np.array(p_ff).T.reshape(2, len(ff)), "fs": np.array(p_fs), "ss": np.array(p_ss), "es": np.array(p_es)}
    return Deck(folder=folder, topology=topo, model=model, chan_ids=chan_ids, solid_ids=solid_ids,
                length=float(_get(cin, "ZLENGTH", "XLENGTH")), n_elems=int(grid["NELEMS"]),
                mesh={k: float(v) for k, v in grid.items() if isinstance(v, float)}, transient=transient,
                t_env=float(env.get("Temperature", 300.0)), fl_bc=fl_bc, solid_ops=solid_ops, peri=peri, notes=notes,
                solid_inputs=[dict(sd[2]) for sd in solids], ss_view_factor=v_ss)
