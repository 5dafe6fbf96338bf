"""Host-side (numpy) restatement of the channel friction laws the device evaluates in K_P
(`csrc/props_kernels.cuh: friction_total`), used only where the reference itself works on the host
outside the time step: the initial flow field (`gen_flow`, utility_functions/gen_flow.py) and the `.tsv`
writers (utility_functions/output.py).  Reference: `Channel.eval_friction_factor`
(/root/reference/source_code/channel.py:1119-1138) with the dispatch of `Channel.__init__` (:43-146):

* -99                      user defined: laminar = turbulent = total = 1 (:248-282)
* 110 / 111 / 112          Blasius (:392-401), Bhatti-Shah (:403-413), LHC recipe (:415-438); laminar slot =
                           16 / Re where Re >= 1e-5 (:284-299); total = max(laminar, turbulent) (:1104-1115)
* 121 / 122 / 124          Haaland (:700-717), Colebrook (:721-756, array-wide iteration count), ENEA HTS
                           CICC hole (:782-795); their "laminar" slot calls the same method, which writes the
                           TURBULENT array, so laminar stays 0 and total = max(0, turbulent)
* 204 / 205, 206 / 209     Katheder and Darcy-Forchheimer bundle laws
Every total is multiplied by FRICTION_MULTIPLIER.
"""
from __future__ import annotations

import numpy as np

SUPPORTED_IFRICTION = (-99, 110, 111, 112, 121, 122, 124, 204, 205, 206, 209)


def _laminar(re: np.ndarray) -> np.ndarray:
    lam = np.zeros(re.shape)
    ind = re >= 1.0e-5
    lam[ind] = 16.0 / re[ind]
    return lam


def _haaland(re, rough, dh):
    return ((-1.8 * np.log10((rough / dh / 3.7) ** 1.11 + (6.9 / re))) ** -2) / 4.0


def friction_factor(ifriction: int, multiplier: float, reynolds, dh: float, roughness: float = 0.0,
                    void_fraction: float = 0.0) -> np.ndarray:
    """Total friction factor (Fanning) of one channel for an array of Reynolds numbers."""
    re = np.fabs(np.atleast_1d(np.asarray(reynolds, dtype=np.float64)))
    flag = int(ifriction)
    if flag == -99:
        total = np.ones(re.shape)
    elif flag == 124:
        f = np.array((0.316 * re ** -0.25) / 4.0)
        f[re > 1e4] = 2.21 * re[re > 1e4] ** (-0.4) / 4.0
        total = np.maximum(np.zeros(re.shape), f)
    elif flag == 121:
        total = np.maximum(np.zeros(re.shape), _haaland(re, roughness, dh))
    elif flag == 122:
        f = _haaland(re, roughness, dh)
        err, it = 10.0 * np.ones(re.shape), 0
        while err.max() > 1e-5 and it < 100:
            fn = ((-2.0 * np.log10((roughness / dh / 3.7) + (2.51 / re / np.sqrt(f)))) ** -2) / 4.0
            err, f, it = abs(f - fn), fn, it + 1
        total = np.maximum(np.zeros(re.shape), f)
    elif flag in (110, 111, 112, 204, 205, 206, 209):
        turb = np.zeros(re.shape)
        if flag == 110:
            turb = np.array((0.316 * re ** -0.25) / 4.0)
        elif flag == 111:
            turb = 0.00128 + 0.1143 * np.maximum(re, 4000) ** -0.311
        elif flag == 112:
            ind = re <= 2.0e3
            turb[ind] = 64 / re[ind] / 4.0
            ind = (re > 2.0e3) & (re < 4.0e3)
            turb[ind] = 2.6471e-3 * re[ind] ** 0.3279 / 4.0
            ind = (re > 4.0e3) & (re < 3.0e4)
            turb[ind] = 0.3194 * re[ind] ** -0.25 / 4.0
            ind = re >= 3.0e4
            turb[ind] = 1.0 / (1.8 * np.log10(re[ind]) / np.log10(10.0) - 1.64) ** 2.0 / 4.0
        elif flag in (204, 205):
            a, c = (0.843, 0.0265) if flag == 204 else (0.88, 0.051)
            turb = void_fraction ** 0.72 * (19.5 / re ** a + c) / 4.0
        else:
            c0, c1, c2 = (19.6e-9, 2.42, 5.80) if flag == 206 else (20.9e-9, 19.1, 4.23)
            jj = c1 / void_fraction ** c2
            kk = c0 * void_fraction ** 3.0 / (1 - void_fraction) ** 2.0
            turb = dh ** 2.0 * void_fraction / 2.0 / kk / re + (dh * void_fraction ** 2.0) / 2.0 * jj
        total = np.maximum(_laminar(re), turb)
    else:
        raise ValueError(f"IFRICTION {flag} is not implemented (device flags: {SUPPORTED_IFRICTION})")
    return total * multiplier


def channel_friction_factor(model, topo, j: int, reynolds) -> np.ndarray:
    """`friction_factor` for channel j of a (Model, Topology) pair."""
    return friction_factor(model.ch_ifriction[j], model.ch_friction_mult[j], reynolds, topo.ch_dh[j],
                           model.ch_roughness[j] if model.ch_roughness else 0.0,
                           model.ch_void_fraction[j] if model.ch_void_fraction else 0.0)
