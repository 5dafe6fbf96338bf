"""Spatial discretisation of a conductor: the `zcoord` array the time-step path takes as an input.

Host-side restatement (runs once per run) of the reference's
`conductor_spatial_discretization` (/root/reference/source_code/utility_functions/initialization_functions.py:22-62):
  * `uniform_mesh`  -- `uniform_spatial_discretization` (:65-76): ITYMSH 0 / 2, or a refined region narrower
    than 1 mm;
  * `refined_mesh`  -- `fixed_refined_spatial_discretization` (:120-275): ITYMSH 1 / 3, a uniformly refined
    window [XBREFI, XEREFI] of NELREF elements, element sizes growing by DXINCRE_LEFT / _RIGHT away from it until
    the remaining span can be covered uniformly.
The legacy decks of TDD_examples carry one `DXINCRE` for both sides.
"""
from __future__ import annotations

from typing import Mapping

import numpy as np


def uniform_mesh(length: float, n_nodes: int) -> np.ndarray:
    return np.linspace(0.0, length, n_nodes)


def refined_mesh(length: float, n_elems: int, n_elref: int, xbrefi: float, xerefi: float, sizmin: float,
                 sizmax: float, dxincre_left: float, dxincre_right: float) -> np.ndarray:
    n_nodes = n_elems + 1
    z = np.zeros(n_nodes)
    n_coarse = n_elems - n_elref
    n_left = round((xbrefi - 0.0) / (length - (xerefi - xbrefi)) * n_coarse)
    n_right = n_coarse - n_left
    nod_ref = n_elref + 1
    dx_ref = (xerefi - xbrefi) / n_elref
    if dx_ref < sizmin:
        raise ValueError(f"ERROR: dx_ref={dx_ref} m in refined zone < SIZMIN={sizmin} m!!!\n")
    if dx_ref > sizmax:
        raise ValueError(f"ERROR: dx_ref={dx_ref} m in refined zone > {sizmax} m!!!\n")
    z[n_left:n_left + nod_ref] = np.linspace(xbrefi, xerefi, nod_ref)

    if n_left > 0:
        dx_try = (xbrefi - 0.0) / n_left
        dx1, ii = dx_ref, 0
        while (dx_try / dx1 > dxincre_left) and (ii < n_left - 1):
            ii += 1
            dx = dx1 * dxincre_left
            z[n_left - ii] = z[n_left + 1 - ii] - dx
            dx1 = dx
            dx_try = (z[n_left - ii] - 0.0) / (n_left - ii)
        if ii < n_left - 1:
            z[:n_left + 1 - ii] = np.linspace(0.0, z[n_left - ii], n_left + 1 - ii)
            if z[n_left - ii] / (n_left - ii) < dx1:
                raise ValueError("Bad spatial discretization.\nDiscretization pitch in used in the uniform region to the "
                                 "left of the refined region is lower than the last discretization pitch used for coarsening.")
        elif z[1] > dx1 * dxincre_left:
            raise ValueError("Bad spatial discretization.\nElement lenght between the second and first node is larger "
                             "than the expected one.")
    if n_right > 0:
        dx_try = (length - xerefi) / n_right
        dx1, ii = dx_ref, 0
        base = n_left + n_elref
        while (dx_try / dx1 > dxincre_right) and (ii < n_right - 1):
            ii += 1
            dx = dx1 * dxincre_right
            z[base + ii] = z[base + ii - 1] + dx
            dx1 = dx
            dx_try = (length - z[base + ii]) / (n_right - ii)
        if ii < n_right - 1:
            z[base + ii:] = np.linspace(z[base + ii], length, n_right + 1 - ii)
            if (length - z[base + ii]) / (n_right - ii) < dx1:
                raise ValueError("Bad spatial discretization.\nDiscretization pitch in used in the uniform region to the "
                                 "right of the refined region is lower than the last discretization pitch used for coarsening.")
        else:
            z[-1] = length
            if z[-1] > z[-2] + dx1 * dxincre_right:
                raise ValueError("Bad spatial discretization.\nElement lenght between the last but one and last node is "
                                 "larger than the expected one.")
    return z


def build_mesh(length: float, grid: Mapping[str, float], n_elems: int | None = None) -> np.ndarray:
    """`zcoord` from the GRID sheet of a deck (`Deck.mesh`).  `n_elems` overrides NELEMS (uniform meshes only:
    the batch configurations refine the CASE_1 mesh to 10 000 elements)."""
    nel = int(grid["NELEMS"]) if n_elems is None else int(n_elems)
    itymsh = int(grid.get("ITYMSH", 0))
    xb, xe = float(grid.get("XBREFI", 0.0)), float(grid.get("XEREFI", 0.0))
    if itymsh in (0, 2) or abs(xe - xb) <= 1e-3:
        z = uniform_mesh(length, nel + 1)
    elif itymsh in (1, 3):
        if n_elems is not None and int(n_elems) != int(grid["NELEMS"]):
            raise ValueError("n_elems override is only defined for uniform meshes")
        inc = grid.get("DXINCRE")
        z = refined_mesh(length, nel, int(grid["NELREF"]), xb, xe, float(grid["SIZMIN"]), float(grid["SIZMAX"]),
                         float(grid.get("DXINCRE_LEFT", inc)), float(grid.get("DXINCRE_RIGHT", inc)))
    else:
        raise ValueError(f"ITYMSH {itymsh} (user defined mesh file) is not supported: pass zcoord directly")
    # check_grid_features (initialization_functions.py:578-): the mesh must span [0, ZLENGTH]
    if z[0] != 0.0 or abs(z[-1] - length) > 1e-12 * max(1.0, length):
        raise ValueError(f"mesh does not span the conductor: [{z[0]}, {z[-1]}] vs ZLENGTH {length}")
    return z
