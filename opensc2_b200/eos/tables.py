"""(p, T) property tables for the device: `props[10][n_p][n_t]` on a regular grid, in the property order of the
reference's alias table (simulation.py:89-100; `model.FP_*`), and the error a BILINEAR lookup in them makes
against the equation of state itself -- the number the grid is chosen from."""
from __future__ import annotations

from typing import Callable, Dict, Tuple

import numpy as np

from ..model import FluidTable

ORDER = ("beta", "kappa", "prandtl", "rho", "mu", "h", "cp", "cv", "w", "k")


def build_table(state: Callable, p_range: Tuple[float, float], t_range: Tuple[float, float], n_p: int, n_t: int) -> FluidTable:
    """`state(T, p)` -> dict of the ten properties (e.g. `eos.nitrogen.state`)."""
    p = np.linspace(p_range[0], p_range[1], n_p)
    t = np.linspace(t_range[0], t_range[1], n_t)
    P, T = np.meshgrid(p, t, indexing="ij")
    s = state(T.ravel(), P.ravel())
    props = np.stack([np.asarray(s[k]).reshape(n_p, n_t) for k in ORDER])
    if not np.all(np.isfinite(props)):
        raise ValueError("the equation of state returned non-finite values inside the table range")
    return FluidTable(float(p[0]), float(p[1] - p[0]), float(t[0]), float(t[1] - t[0]), props)


def interpolation_error(table: FluidTable, state: Callable, n_probe: int = 4000, seed: int = 0) -> Dict[str, float]:
    """max relative error of the bilinear lookup against `state` at random interior points."""
    from ..model import table_lookup

    rng = np.random.default_rng(seed)
    n_p, n_t = table.props.shape[1:]
    p = table.p_min + rng.uniform(0.0, (n_p - 1) * table.p_step, n_probe)
    t = table.t_min + rng.uniform(0.0, (n_t - 1) * table.t_step, n_probe)
    s = state(t, p)
    return {k: float(np.max(np.abs(table_lookup(table, i, p, t) / s[k] - 1.0))) for i, k in enumerate(ORDER)}
