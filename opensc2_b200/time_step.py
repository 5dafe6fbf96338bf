"""Time-step selection, host side: drop-in for the reference's `get_time_step`
(/root/reference/source_code/utility_functions/transient_solution_functions.py:38-184).

`get_time_step(conductor, transient_input, num_step)` keeps the reference's signature and
mutates `conductor.time_step`.  `next_time_step(...)` is the same rule on plain numbers, used for
batches (a batch shares one dt: the smallest of its conductors' steps).

Quirks kept on purpose: STPMAX is never used; the step grows by 1.2 when it is below half the
optimum and halves when above it; the lower clamp is STPMIN; EQTEIG is the device's (quirky)
eigenvalue estimate of step()."""
from __future__ import annotations

import numpy as np

TINY = 1.0e-10
FACTUP = 1.2
FACTLO = 0.5


def _schedule_mode3(t: float, shift: float, tend: float) -> float:
    """IADAPTIME == 3: the ad-hoc schedule of tsf:55-83."""
    times = np.array([0.0, shift, shift + 10.0 - 1.0, shift + 69.0, shift + 69.8, shift + 69.9, shift + 70.0,
                      shift + 70.1, shift + 70.15, shift + 70.5, shift + 71.0, shift + 73.0])
    steps = [1.0, 1e-1, 1e-2, 1e-3, 1e-4, 1e-5, 1e-4, 1e-3, 5e-3, 1e-2]
    if times[0] <= t <= times[1]:
        return steps[0]
    for i in range(1, 10):
        if times[i] <= t < times[i + 1]:
            return steps[i]
    return min(5e-2, tend - t)


def next_time_step(prev_step: float, eqteig, n_fluid: int, n_solid: int, time: float, num_step: int,
                   stpmin: float, tend: float, iadaptime: int = 0, eigtim: float = 1.0, time_shift: float = 0.0) -> float:
    if num_step == 1:
        return stpmin
    if iadaptime == 0:
        return min(stpmin, tend - time)
    if iadaptime == 3:
        return _schedule_mode3(time, time_shift, tend)
    eq = np.asarray(eqteig, dtype=np.float64)
    t_comp = np.zeros(3 * n_fluid + n_solid)
    for ii in range(n_fluid):
        if abs(iadaptime) == 1:
            t_comp[ii] = eigtim / (eq[ii] + TINY)
            t_comp[ii + n_fluid] = eigtim / (eq[ii + n_fluid] + TINY)
        elif iadaptime == 2:
            t_comp[ii] = 1.0e10
            t_comp[ii + n_fluid] = 1.0e10
        t_comp[ii + 2 * n_fluid] = eigtim / (eq[ii + 2 * n_fluid] + TINY)
    for ii in range(n_solid):
        t_comp[ii + 3 * n_fluid] = eigtim / (eq[ii + 3 * n_fluid] + TINY)
    optstp = min(t_comp)
    step = prev_step
    if step < 0.5 * optstp:
        step = step * FACTUP
    elif step > 1.0 * optstp:
        step = step * FACTLO
    step = max(step, stpmin)
    return min(step, tend - time)


def get_time_step(conductor, transient_input, num_step):
    """Reference-compatible entry point (returns None, sets `conductor.time_step`)."""
    conductor.time_step = next_time_step(
        getattr(conductor, "time_step", transient_input["STPMIN"]), getattr(conductor, "EQTEIG", None),
        conductor.inventory["FluidComponent"].number, conductor.inventory["SolidComponent"].number,
        conductor.cond_time[-1], num_step, transient_input["STPMIN"], transient_input["TEND"],
        transient_input["IADAPTIME"], getattr(conductor, "EIGTIM", 1.0), transient_input.get("TIME_SHIFT", 0.0))


def batch_time_step(prev_step: float, eqteig: np.ndarray, n_fluid: int, n_solid: int, time: float, num_step: int,
                    stpmin: float, tend: float, iadaptime: int = 0, eigtim: float = 1.0) -> float:
    """One dt for a whole batch: the smallest step any of its conductors would take (`eqteig`: [B][d])."""
    return min(next_time_step(prev_step, e, n_fluid, n_solid, time, num_step, stpmin, tend, iadaptime, eigtim)
               for e in np.atleast_2d(eqteig))
