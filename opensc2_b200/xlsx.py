"""Minimal .xlsx reader on the standard library (zipfile + xml): the reference parses its input
decks with openpyxl / pandas.read_excel (conductor.py:99-306, 680-863), neither usable in this
image.  Returns every sheet as a dense list of rows (None for empty cells); numbers as float,
shared and inline strings as str, booleans as bool."""
from __future__ import annotations

import re
import zipfile
from typing import Dict, List
from xml.etree import ElementTree as ET

_NS = {"m": "http://schemas.openxmlformats.org/spreadsheetml/2006/main",
       "r": "http://schemas.openxmlformats.org/officeDocument/2006/relationships",
       "p": "http://schemas.openxmlformats.org/package/2006/relationships"}


def _col_index(ref: str) -> int:
    n = 0
    for ch in re.match(r"[A-Z]+", ref).group(0):
        n = n * 26 + (ord(ch) - 64)
    return n - 1


def read_workbook(path: str) -> Dict[str, List[list]]:
    with zipfile.ZipFile(path) as z:
        shared = []
        if "xl/sharedStrings.xml" in z.namelist():
            root = ET.fromstring(z.read("xl/sharedStrings.xml"))
            for si in root.findall("m:si", _NS):
                shared.append("".join(t.text or "" for t in si.iter("{%s}t" % _NS["m"])))
        rels = {r.get("Id"): r.get("Target") for r in ET.fromstring(z.read("xl/_rels/workbook.xml.rels")).findall("p:Relationship", _NS)}
        wb = ET.fromstring(z.read("xl/workbook.xml"))
        sheets = {}
        for s in wb.find("m:sheets", _NS).findall("m:sheet", _NS):
            target = rels[s.get("{%s}id" % _NS["r"])]
            target = target.lstrip("/")
            if not target.startswith("xl/"):
                target = "xl/" + target
            rows: List[list] = []
            root = ET.fromstring(z.read(target))
            for row in root.iter("{%s}row" % _NS["m"]):
                r_idx = int(row.get("r")) - 1
                while len(rows) <= r_idx:
                    rows.append([])
                cur = rows[r_idx]
                for c in row.findall("m:c", _NS):
                    ci = _col_index(c.get("r"))
                    while len(cur) <= ci:
                        cur.append(None)
                    t, v = c.get("t"), c.find("m:v", _NS)
                    if t == "inlineStr":
                        cur[ci] = "".join(x.text or "" for x in c.iter("{%s}t" % _NS["m"]))
                    elif v is None or v.text is None:
                        cur[ci] = None
                    elif t == "s":
                        cur[ci] = shared[int(v.text)]
                    elif t == "b":
                        cur[ci] = bool(int(v.text))
                    elif t in ("str", "e"):
                        cur[ci] = v.text
                    else:
                        cur[ci] = float(v.text)
            sheets[s.get("name")] = rows
        return sheets


def variable_table(rows: List[list], header_key: str = "Variable name") -> Dict[str, Dict[str, object]]:
    """The decks' input sheets: a header row `... | Variable name | <object 1> | <object 2> ...` followed by
    one row per variable.  Returns {object: {variable: value}}."""
    for i, row in enumerate(rows):
        if header_key in row:
            k = row.index(header_key)
            names = [(j, row[j]) for j in range(k + 1, len(row)) if isinstance(row[j], str)]
            out: Dict[str, Dict[str, object]] = {n: {} for _, n in names}
            for r in rows[i + 1:]:
                if len(r) <= k or not isinstance(r[k], str):
                    continue
                for j, n in names:
                    out[n][r[k]] = r[j] if j < len(r) else None
            return out
    return {}


def matrix_table(rows: List[list]) -> Dict[str, Dict[str, object]]:
    """The coupling sheets: a square table whose first header row and first column hold the component
    identifiers.  Returns {row id: {column id: value}}."""
    for i, row in enumerate(rows):
        labels = [(j, v) for j, v in enumerate(row) if isinstance(v, str)]
        if len(labels) >= 2 and i + 1 < len(rows):
            cols = labels
            out: Dict[str, Dict[str, object]] = {}
            for r in rows[i + 1:]:
                ids = [v for v in r if isinstance(v, str)]
                if not ids:
                    continue
                rid = ids[0]
                k = r.index(rid)
                out[rid] = {name: (r[j] if j < len(r) else None) for j, name in cols if j > k}
            if out:
                return out
    return {}
