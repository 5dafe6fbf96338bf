"""tests/golden/decks_summary.json: what `opensc2_b200.deck.load_deck` extracts from the shipped
TDD_examples decks (plain numbers, no reference code).  Build-container only."""
import dataclasses
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from opensc2_b200.deck import deck_to_dict, load_deck  # noqa: E402

CASES = {"CASE_1_ITER_like_LTS": 1, "CASE_2_ENEA_HTS_CICC": 3, "CASE_3_HTS_HVDC": 3}


def summary(d):
    return deck_to_dict(d)


def main():
    out = {c: summary(load_deck(os.path.join("/root/reference/TDD_examples", c), intial_override=ov)) for c, ov in CASES.items()}
    with open(os.path.join(HERE, "decks_summary.json"), "w") as f:
        json.dump(out, f, indent=1)
    print({c: (v["topology"]["n_ch"], v["topology"]["n_sol"]) for c, v in out.items()})


if __name__ == "__main__":
    main()
