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
* 113 / 114 / 115 / 118    hole laws (:447-496, 636-648; the closing statement of 113 has no effect)
* 200-203, 207, 210        bundle laws (:799-866, 939-963, 1015-1037)
* 211                      HTS current-lead bundle (:1039-1066): laminar slot stays 0, FRICTION_MULTIPLIER forced to 1
* 100 / 101                ITER spiral holes (:321-347)
* 102-109                  hole laws by fixed-point iteration on the whole array (:190-222, 349-390); 106 interpolates
                           between laminar and turbulent (:1068-1101), the others take the maximum
* 119 / 120                rectangular (SIDE2 / SIDE1) and triangular duct (:650-698), total by interpolation
Every total is multiplied by FRICTION_MULTIPLIER.  Not built because they do not run in the reference: 116 / 117 (`dict_coeff(flag_eval_ccc)` calls a
dict: TypeError, channel.py:498-555), 123 (`and` on arrays, a misspelt method: :759-780), 208 (raises ValueError, :965-987).
"""
from __future__ import annotations

import numpy as np

SUPPORTED_IFRICTION = (-99, 100, 101, 102, 103, 104, 105, 106, 107, 108, 109, 110, 111, 112, 113, 114, 115, 118, 119, 120, 121, 122, 124,
                       200, 201, 202, 203, 204, 205, 206, 207, 209, 210, 211)


def _laminar(re: np.ndarray) -> np.ndarray:
    lam = np.zeros(re.shape)
    ind = re >= 1.0e-5
    lam[ind] = 16.0 / re[ind]
    return lam


def _total_llimt(re, lam, turb, l_lim_t):
    """_eval_total_friction_factor_L_lim_T (channel.py:1068-1101); l_lim_t is what the caller really passes (see friction_factor)."""
    tot = np.zeros(re.shape)
    ind = re <= 2.0e3
    tot[ind] = lam[ind]
    ind = (re > 2.0e3) & (re < l_lim_t)
    tot[ind] = (re[ind] - 2000.0) / (l_lim_t - 2000.0) * (turb[ind] - lam[ind]) + lam[ind]
    ind = re >= l_lim_t
    tot[ind] = turb[ind]
    return tot


def _haaland(re, rough, dh):
    return ((-1.8 * np.log10((rough / dh / 3.7) ** 1.11 + (6.9 / re))) ** -2) / 4.0


def friction_factor(ifriction: int, multiplier: float, reynolds, dh: float, roughness: float = 0.0,
                    void_fraction: float = 0.0, area: float = 0.0, nodal=True, side_ratio: float = 0.0) -> np.ndarray:
    """Total friction factor (Fanning) of one channel for an array of Reynolds numbers.

    nodal: the reference's `nodal` argument (True: nodal pass, False: Gauss points, None: gen_flow).  It only matters
    for the interpolated totals (106, 120): `eval_friction_factor` passes it positionally into the L_lim_T parameter of
    `_eval_total_friction_factor_L_lim_T` (channel.py:1068-1138), so L_lim_T is 1, 0 or -- in gen_flow -- None, where the
    reference raises TypeError."""
    if int(ifriction) in (106, 119, 120) and nodal is None:
        raise TypeError(f"IFRICTION {int(ifriction)} in gen_flow: '>=' not supported between instances of 'float' and "
                        "'NoneType' (the reference compares the Reynolds number with L_lim_T = nodal = None)")
    llim = 1.0 if nodal else 0.0
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
    elif flag in (100, 101):
        aa, bb = (0.45, -0.034) if flag == 100 else (0.36, -0.038)
        corr = (dh + 2 * 1e-3) / dh
        total = np.maximum(_laminar(re), (aa * (re / corr) ** bb) / 4.0 / (corr ** 5.0))
    elif 102 <= flag <= 109:
        fit64 = flag in (102, 106, 107)
        aa, bb, cc = (6.4, 0.1717, -0.3428) if fit64 else (11.88, 0.039, -0.299)
        gg = {103: 3.0e-3, 104: 1.5e-3, 106: 2.0e-3, 108: 2.40e-3, 109: 5.30e-3}.get(flag, 3.05e-3)
        tt = 1.5e-3 if flag == 104 else 1e-3
        goverh, hp, hd2 = gg / tt, tt / dh * re, 2.0 * tt / dh
        turb = 1e-2 * np.ones(re.shape)
        err, it = 10 * np.ones(re.shape), 0
        with np.errstate(all="ignore"):
            while np.max(err) >= 1e-2 and it < 1000:
                old = turb
                hpiu = hp * np.sqrt(0.5 * turb)
                rhpiu = aa * hpiu ** bb * goverh ** cc
                turb = 2.0 / (rhpiu - 2.5 * np.log10(hd2) - 3.75) ** 2
                err, it = np.fabs((turb - old) / turb), it + 1
        total = _total_llimt(re, _laminar(re), turb, llim) if flag == 106 else np.maximum(_laminar(re), turb)
    elif flag in (119, 120):
        from .model import duct_laminar_coefficient
        lamc = duct_laminar_coefficient(side_ratio) if flag == 119 else 13.33
        with np.errstate(all="ignore"):
            total = _total_llimt(re, lamc / np.minimum(re, 2000), 0.00128 + 0.1143 * np.maximum(re, 4000) ** -0.311, llim)
    elif flag in (113, 114, 115, 118, 200, 201, 202, 203, 207, 210, 211):
        turb = np.zeros(re.shape)
        if flag == 113:
            corr = (dh + 2 * 1e-3) / dh
            ind = re / corr < 1.5e5
            turb[ind] = 3.160 * (re[ind] / corr) ** -0.249
            ind = re / corr >= 1.5e5
            turb[ind] = 0.216 * (re[ind] / corr) ** -0.024
        elif flag == 114:
            turb = 0.0958 * re ** -0.181
        elif flag == 115:
            corr = 3.18e-3 / dh
            turb = 0.25 * 0.3164 * (re / corr) ** -0.25 / corr
        elif flag == 118:
            turb = 0.1687 / 4.0 * re ** -0.1129
        elif flag == 200:
            doutct = dh + 2 * 1e-3
            turb = (2.46 * (re / dh * 0.25e-3) ** -0.52 * dh / 0.25e-3
                    * (area / (0.3 * (39.8e-3 ** 2 - doutct ** 2) * 0.25 * np.pi)) ** 2 / 4.0)
        elif flag == 201:
            turb = (0.0231 + 19.5 / (re / (6.0 / 5.0)) ** 0.7953) / (void_fraction ** 0.742) / 4.0
        elif flag == 202:
            turb = 6.1 * re ** -0.51 / 4.0
        elif flag == 203:
            turb = 0.4335 * re ** -0.263
        elif flag == 207:
            ind = (re >= 2e3) & (re < 4e3)
            turb[ind] = 2.6471e-3 * re[ind] ** 0.3279 / 4.0
            ind = (re >= 4e3) & (re < 3e4)
            turb[ind] = 0.3194 * re[ind] ** -0.25 / 4.0
            ind = re >= 3e4
            turb[ind] = 1.0 / (1.8 * np.log10(re[ind]) / np.log10(10.0) - 1.64) ** 2 / 4.0
        elif flag == 210:
            ind = re < 1.5e3
            turb[ind] = 47.65 * re[ind] ** -0.885 / 4.0
            ind = (re >= 1.5e3) & (re <= 2e5)
            turb[ind] = 1.093 * re[ind] ** -0.338 / 4.0
            turb[re > 2e5] = 0.0377 / 4.0
        else:
            ind = re < 1e3
            turb[ind] = 194.002 * re[ind] ** -0.52 / 4.0
            ind = (re >= 1e3) & (re < 2e3)
            turb[ind] = 4.8563 + 4.87e-4 * re[ind] / 4.0
            ind = re >= 2e3
            turb[ind] = 14.5148 * re[ind] ** -0.12 / 4.0
        total = np.maximum(np.zeros(re.shape) if flag == 211 else _laminar(re), turb)
        if flag == 211:
            multiplier = 1.0
    else:
        raise ValueError(f"IFRICTION {flag} is not implemented (device flags: {SUPPORTED_IFRICTION})")
    return total * multiplier


def channel_friction_factor(model, topo, j: int, reynolds, nodal=True) -> np.ndarray:
    """`friction_factor` for channel j of a (Model, Topology) pair."""
    return friction_factor(model.ch_ifriction[j], model.ch_friction_mult[j], reynolds, topo.ch_dh[j],
                           model.ch_roughness[j] if model.ch_roughness else 0.0,
                           model.ch_void_fraction[j] if model.ch_void_fraction else 0.0, topo.ch_area[j], nodal,
                           model.ch_side_ratio[j] if getattr(model, "ch_side_ratio", ()) else 0.0)
