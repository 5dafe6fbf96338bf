"""tests/golden/tdd_solutions.npz: the nodal fields of the reference's own shipped results
(TDD_examples/CASE_*/Solution/*.tsv, written by the reference's save_properties at TEND): per channel z, p, T, rho, v,
mass flow; per solid z, T.  They are the only end-to-end golden vectors the reference has; the GPU box has
no /root/reference, hence this copy of the needed columns.  Build container only."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = "/root/reference/TDD_examples"
CASES = {"case1": "CASE_1_ITER_like_LTS", "case2": "CASE_2_ENEA_HTS_CICC", "case3": "CASE_3_HTS_HVDC"}


def main():
    out = {}
    for tag, folder in CASES.items():
        sol = os.path.join(ROOT, folder, "Solution")
        for f in sorted(os.listdir(sol)):
            a = np.loadtxt(os.path.join(sol, f), skiprows=1)
            name = f[:-4]
            if name.startswith("CHAN"):
                out[f"{tag}/{name}"] = np.stack([a[:, 0], a[:, 1], a[:, 2], a[:, 3], a[:, 4], a[:, -1]], axis=1)   # z p T rho v mdot
                # every column, for the property checks and the local helium table of tests/transient_common.py.  Legacy
                # layout (CASE_1 / CASE_2, 15 columns; the header's unit labels are shuffled, the order below is the real one):
                # x p T rho v mu Gruneisen h cp cv w k Re Pr mdot; HEAD layout (CASE_3, 17 columns): see the file's header
                out[f"{tag}/{name}/all"] = a
            else:
                out[f"{tag}/{name}"] = a[:, :2]                                                                    # z T
    np.savez_compressed(os.path.join(HERE, "tdd_solutions.npz"), **out)
    print(len(out), "components")


if __name__ == "__main__":
    main()
