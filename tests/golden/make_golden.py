"""Generate tests/golden/step_<case>.npz by running the UNMODIFIED reference.

Build-container only (needs /root/reference):  python tests/golden/make_golden.py
Each fixture holds the inputs (so a numpy RNG change cannot silently shift
them) and the outputs of the reference's own step()
(/root/reference/source_code/utility_functions/transient_solution_functions.py:186).
"""
import dataclasses
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import ref_harness as H  # noqa: E402
from cases import STEP_CASES, WITH_SYSTEM, make_case  # noqa: E402


def main():
    assert H.reference_available(), "needs /root/reference"
    for name in STEP_CASES:
        topo, arr = make_case(name)
        ref = H.run_reference_step(topo, arr)
        payload = {}
        for f in dataclasses.fields(arr):
            if getattr(arr, f.name) is not None:
                payload["in_" + f.name] = np.asarray(getattr(arr, f.name))
        for k, v in ref.items():
            if k in ("sysmat", "known") and name not in WITH_SYSTEM:
                continue
            payload["out_" + k] = v
        path = os.path.join(HERE, f"step_{name}.npz")
        np.savez_compressed(path, **payload)
        print(name, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
