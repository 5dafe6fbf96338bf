"""tests/golden/decks_summary.json: what `opensc2_b200.deck.load_deck` extracts from the shipped
TDD_examples decks (plain numbers, no reference code).  Build-container only."""
import dataclasses
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from opensc2_b200.deck import load_deck  # noqa: E402

CASES = {"CASE_1_ITER_like_LTS": 1, "CASE_2_ENEA_HTS_CICC": 3, "CASE_3_HTS_HVDC": 3}


def summary(d):
    return {"topology": dataclasses.asdict(d.topology), "model": dataclasses.asdict(d.model), "chan_ids": d.chan_ids,
            "solid_ids": d.solid_ids, "length": d.length, "n_elems": d.n_elems, "t_env": d.t_env,
            "fl_bc": d.fl_bc.tolist(), "solid_ops": d.solid_ops, "peri": {k: v.tolist() for k, v in d.peri.items()},
            "transient": {k: v for k, v in d.transient.items() if k in ("TEND", "STPMIN", "IADAPTIME")}, "notes": d.notes}


def main():
    out = {c: summary(load_deck(os.path.join("/root/reference/TDD_examples", c), intial_override=ov)) for c, ov in CASES.items()}
    with open(os.path.join(HERE, "decks_summary.json"), "w") as f:
        json.dump(out, f, indent=1, default=lambda x: "inf" if x == float("inf") else x)
    print({c: (v["topology"]["n_ch"], v["topology"]["n_sol"]) for c, v in out.items()})


if __name__ == "__main__":
    main()
