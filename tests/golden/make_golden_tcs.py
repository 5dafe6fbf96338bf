"""tests/golden/tcs_nb3sn.npz: current sharing temperature, critical current density and critical temperature of Nb3Sn
by the UNMODIFIED reference functions (properties_of_materials/niobium3_tin.py: current_sharing_temperature_nb3sn with
its scipy.optimize.bisect(xtol=1e-5), critical_current_density_nb3sn, critical_temperature_nb3sn) on swept field,
strain and operating current density -- the pin of the device kernel osc2_tcs_kernel.  Build container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402


def main():
    H._import_reference()
    from properties_of_materials.niobium3_tin import (critical_current_density_nb3sn, critical_temperature_nb3sn,
                                                       current_sharing_temperature_nb3sn)

    tc0m, bc20m, c0 = 16.22, 32.35, 1.21508e11            # CASE_1 STR_MIX_1
    rng = np.random.default_rng(20261020)
    batch, n = 24, 33                                      # 24 conductors x 33 nodes: BISS..BOSS linspace per conductor
    biss = rng.uniform(0.0, 13.0, batch)
    boss = biss * rng.uniform(0.85, 1.15, batch)
    eps = rng.uniform(-0.007, 0.0, batch)
    jop = rng.uniform(0.0, 2.8e8, batch)
    jop[0] = 0.0                                           # no current: Tcs = Tc
    biss[1], boss[1] = 0.0, 0.0                            # B below BLOW
    jop[2] = 5.0e9                                         # above Jc(T = 0): Tcs = 0
    biss[3], boss[3] = 29.0, 34.0                          # crosses Bc20m * S: Tcs = 0 where Bstar >= 1
    jop[3] = 1.0e6
    tcs_n, tcs_g, tc_g, jc_g = (np.zeros((batch, n)), np.zeros((batch, n - 1)), np.zeros((batch, n - 1)), np.zeros((batch, n - 1)))
    raises = np.zeros(batch, dtype=np.int32)      # the reference's bisect raised "f(a) and f(b) must have different signs"
    for b in range(batch):
        bn = np.linspace(biss[b], boss[b], n)
        bg = np.abs(bn[:-1] + bn[1:]) / 2.0
        eg = np.full(n - 1, (eps[b] + eps[b]) / 2.0)
        try:
            tcs_n[b] = current_sharing_temperature_nb3sn(bn, np.full(n, eps[b]), np.full(n, jop[b]), tc0m, bc20m, c0)
            tcs_g[b] = current_sharing_temperature_nb3sn(bg, eg, np.full(n - 1, jop[b]), tc0m, bc20m, c0)
        except ValueError as e:
            assert "different signs" in str(e)
            raises[b] = 1
            tcs_n[b], tcs_g[b] = 0.0, 0.0
        tc_g[b] = critical_temperature_nb3sn(bg, eg, tc0m, bc20m)
        jc_g[b] = critical_current_density_nb3sn(np.full(n - 1, 4.5), bg, eg, tc0m, bc20m, c0)
    np.savez_compressed(os.path.join(HERE, "tcs_nb3sn.npz"), biss=biss, boss=boss, eps=eps, jop=jop, tcs_nodal=tcs_n, tcs_gauss=tcs_g,
                        tc_gauss=tc_g, jc_gauss_4p5=jc_g, raises=raises, tc0m=tc0m, bc20m=bc20m, c0=c0)
    print("conductors on which the reference raises:", np.nonzero(raises)[0])
    print("tcs range", tcs_g.min(), tcs_g.max(), "zeros:", int((tcs_g == 0).sum()))


if __name__ == "__main__":
    main()
