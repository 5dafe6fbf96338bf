"""Batched conductor step on one B200: Python face of the C ABI.

`BatchedStep` is the batch counterpart of the reference's
`step(conductor, environment, qsource, num_step)`
(/root/reference/source_code/utility_functions/transient_solution_functions.py:186):
one call advances B independent conductors that share a topology.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _lib
from ._abi import (KERNEL_NAMES, OSC2_K_COUNT, STEP_INPUT_FIELDS, Osc2StepInputs, as_f64, c_ip,
                   dptr, topology_struct)
from .problem import StepArrays, Topology


class BatchedStep:
    def __init__(self, topo: Topology, batch: int, n_nodes: int, device: int = 0):
        self._L = _lib.load()
        self.topo = topo
        self.batch, self.n_nodes = int(batch), int(n_nodes)
        self.ndf = topo.ndf
        self.total = self.ndf * self.n_nodes
        self._h = C.c_void_p()
        self._t_env = 300.0
        self._tstruct = topology_struct(topo)
        _lib.check(self._L.osc2_create(C.byref(self._tstruct), self.batch, self.n_nodes, int(device), C.byref(self._h)))

    # -- lifetime -----------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.osc2_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- data movement ------------------------------------------------------
    def _shape(self, a, shape, name):
        a = as_f64(a)
        if a.shape != shape:
            raise ValueError(f"{name}: expected shape {shape}, got {a.shape}")
        return a

    def upload_state(self, sysvar=None, syslod=None):
        sv = self._shape(sysvar, (self.batch, self.total), "sysvar") if sysvar is not None else None
        sl = self._shape(syslod, (self.batch, self.total, 2), "syslod") if syslod is not None else None
        _lib.check(self._L.osc2_upload_state(self._h, dptr(sv) if sv is not None else None,
                                             dptr(sl) if sl is not None else None))

    def upload_step_inputs(self, arr: StepArrays):
        s = Osc2StepInputs()
        keep = []
        for name in STEP_INPUT_FIELDS:
            a = as_f64(getattr(arr, name))
            if a.shape[0] != self.batch:
                raise ValueError(f"{name}: leading axis {a.shape[0]} is not the batch {self.batch}")
            keep.append(a)
            setattr(s, name, dptr(a) if a.size else None)
        s.t_env = self._t_env = float(arr.t_env)
        _lib.check(self._L.osc2_upload_step_inputs(self._h, C.byref(s)))

    # -- device-side operating conditions (K_P) ---------------------------------
    def set_model(self, model, tables=()):
        """`model`: opensc2_b200.model.Model; `tables`: FluidTable per fluid index."""
        self._model_struct = model.struct(self.topo)
        _lib.check(self._L.osc2_set_model(self._h, C.byref(self._model_struct)))
        for i, t in enumerate(tables):
            _lib.check(self._L.osc2_set_fluid_table(self._h, i, t.props.shape[1], t.props.shape[2], t.p_min, t.p_step,
                                                    t.t_min, t.t_step, dptr(t.props)))

    def upload_sweep(self, sweep):
        s, keep = sweep.struct()
        _lib.check(self._L.osc2_upload_sweep(self._h, C.byref(s)))
        del keep

    def upload_geometry(self, arr: StepArrays):
        """Only what K_P does not compute: dz, contact perimeters, qsource, boundary values."""
        s = Osc2StepInputs()
        keep = []
        for name in ("dz", "fl_bc", "peri_ff", "peri_fs", "peri_ss", "peri_es", "qsource", "peri_ss_nodal"):
            v = getattr(arr, name)
            if v is None:
                continue
            a = as_f64(v)
            keep.append(a)
            setattr(s, name, dptr(a) if a.size else None)
        s.t_env = self._t_env = float(arr.t_env)
        _lib.check(self._L.osc2_upload_step_inputs(self._h, C.byref(s)))

    def eval_operating_conditions(self, time: float):
        _lib.check(self._L.osc2_eval_operating_conditions(self._h, float(time)))

    def eval_tcs(self):
        """Current sharing temperature of the Nb3Sn strands on the device, once at step 0 (osc2_eval_tcs)."""
        _lib.check(self._L.osc2_eval_tcs(self._h))

    def download_tcs(self, nodal: bool = True):
        """(tcs_gauss (B, S, E), tcs_nodal (B, S, N) or None)."""
        S = self.topo.n_sol
        g = np.empty((self.batch, S, self.n_nodes - 1))
        nd = np.empty((self.batch, S, self.n_nodes)) if nodal else None
        _lib.check(self._L.osc2_download_tcs(self._h, dptr(g), dptr(nd) if nd is not None else None))
        return g, nd

    def download_margin(self) -> np.ndarray:
        """min over the Gauss points of (Tcs - T), (B, S), as of the last operating-condition evaluation."""
        m = np.empty((self.batch, self.topo.n_sol))
        _lib.check(self._L.osc2_download_margin(self._h, dptr(m)))
        return m

    NODAL_COLUMNS = ("total_density", "isobaric_expansion_coefficient", "isothermal_compressibility", "Prandtl",
                     "total_dynamic_viscosity", "total_enthalpy", "total_isobaric_specific_heat",
                     "total_isochoric_specific_heat", "total_speed_of_sound", "total_thermal_conductivity", "Reynolds",
                     "Gruneisen", "mass_flow_rate", "friction_factor")

    def nodal_properties(self) -> Dict[str, np.ndarray]:
        """Nodal coolant pass on the current device state (coolant.py:134-263 with nodal = True, post-step density and
        mass flow rate, nodal friction factor): {column name: (B, M, N)}, the property columns of save_properties'
        channel files."""
        _lib.check(self._L.osc2_eval_nodal_properties(self._h))
        a = np.empty((len(self.NODAL_COLUMNS), self.batch, self.topo.n_ch, self.n_nodes))
        _lib.check(self._L.osc2_download_nodal_properties(self._h, dptr(a)))
        return {k: a[i] for i, k in enumerate(self.NODAL_COLUMNS)}

    def advance(self, time: float, dt: float, num_step: int):
        """Operating conditions at `time` + one step()."""
        _lib.check(self._L.osc2_advance(self._h, float(time), float(dt), int(num_step)))

    def download_step_inputs(self) -> Dict[str, np.ndarray]:
        B, E, M, S = self.batch, self.n_nodes - 1, self.topo.n_ch, self.topo.n_sol
        t = self.topo
        out = {"fl_gauss": np.empty((B, 9, M, E)), "fl_node": np.empty((B, 4, M)), "sol_gauss": np.empty((B, 3, S, E)),
               "sol_q": np.empty((B, 2, S, E, 2)), "htc_ff": np.empty((B, 2, len(t.ff), E)),
               "htc_fs": np.empty((B, len(t.fs), E)), "htc_ss": np.empty((B, len(t.ss), E)),
               "htc_es": np.empty((B, len(t.es), 3, E))}
        ptrs = [dptr(out[k]) if out[k].size else None for k in
                ("fl_gauss", "fl_node", "sol_gauss", "sol_q", "htc_ff", "htc_fs", "htc_ss", "htc_es")]
        _lib.check(self._L.osc2_download_step_inputs(self._h, *ptrs))
        return out

    def step(self, dt: float, num_step: int):
        _lib.check(self._L.osc2_step(self._h, float(dt), int(num_step)))

    def synchronize(self):
        _lib.check(self._L.osc2_synchronize(self._h))

    def download_state(self, want_syslod: bool = True):
        sv = np.empty((self.batch, self.total))
        sl = np.empty((self.batch, self.total, 2)) if want_syslod else None
        st = np.empty(self.batch, dtype=np.int32)
        _lib.check(self._L.osc2_download_state(self._h, dptr(sv), dptr(sl) if sl is not None else None,
                                               st.ctypes.data_as(c_ip)))
        return sv, sl, st

    def download_diagnostics(self, out: Optional[Dict[str, np.ndarray]] = None, want_k123: bool = True) -> Dict[str, np.ndarray]:
        """`out`: optional preallocated (e.g. page-locked) arrays to fill."""
        d, E, nff = self.ndf, self.n_nodes - 1, len(self.topo.ff)
        if out is None:
            out = {"norm_solution": np.empty((self.batch, d)), "norm_change": np.empty((self.batch, d)),
                   "eqteig": np.empty((self.batch, d))}
            if want_k123:
                out["k123"] = np.empty((self.batch, 3, nff, E))
        k123 = out.get("k123")
        _lib.check(self._L.osc2_download_diagnostics(
            self._h, dptr(out["norm_solution"]), dptr(out["norm_change"]), dptr(out["eqteig"]),
            dptr(k123) if k123 is not None and k123.size else None))
        return out

    def upload_boundary_values(self, fl_bc):
        """Per-step boundary values only (PREINL, PREOUT, TEMINL, TEMOUT, MDTIN, MDTOUT): [B][6][M]."""
        a = self._shape(fl_bc, (self.batch, 6, self.topo.n_ch), "fl_bc")
        s = Osc2StepInputs()
        s.fl_bc = dptr(a)
        s.t_env = self._t_env          # the struct carries the environment temperature too
        _lib.check(self._L.osc2_upload_step_inputs(self._h, C.byref(s)))

    def download_status(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        st = out if out is not None else np.empty(self.batch, dtype=np.int32)
        _lib.check(self._L.osc2_download_state(self._h, None, None, st.ctypes.data_as(c_ip)))
        return st

    def set_capture(self, on: bool):
        _lib.check(self._L.osc2_set_capture(self._h, int(on)))

    def download_system(self):
        d = self.ndf
        sysmat = np.empty((self.batch, 4 * d - 1, self.total))
        known = np.empty((self.batch, self.total))
        _lib.check(self._L.osc2_download_system(self._h, dptr(sysmat), dptr(known)))
        return sysmat, known

    # -- timing -------------------------------------------------------------
    def set_profiling(self, on: bool):
        _lib.check(self._L.osc2_set_profiling(self._h, int(on)))

    def kernel_ms(self, reset: bool = False):
        ms = (C.c_double * OSC2_K_COUNT)()
        n = (C.c_longlong * OSC2_K_COUNT)()
        _lib.check(self._L.osc2_get_kernel_ms(self._h, C.byref(ms), C.byref(n)))
        out = {KERNEL_NAMES[k]: (ms[k], n[k]) for k in range(OSC2_K_COUNT) if n[k]}
        if reset:
            _lib.check(self._L.osc2_reset_kernel_ms(self._h))
        return out

    def timer_start(self):
        _lib.check(self._L.osc2_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        _lib.check(self._L.osc2_timer_stop(self._h, C.byref(ms)))
        return ms.value

    @property
    def device_bytes(self) -> int:
        return int(self._L.osc2_device_bytes(self._h))

    # -- convenience: one full reference-equivalent step on host arrays -------
    def run(self, arr: StepArrays, capture: bool = False) -> Dict[str, np.ndarray]:
        """upload -> step -> download; `arr` is batched.  Raises on a singular system."""
        if capture:
            self.set_capture(True)
        self.upload_state(arr.sysvar, arr.syslod)
        self.upload_step_inputs(arr)
        self.step(arr.dt, arr.num_step)
        self.synchronize()
        sv, sl, st = self.download_state()
        out = {"sysvar": sv, "syslod": sl, "status": st}
        out.update(self.download_diagnostics())
        if capture:
            out["sysmat"], out["known"] = self.download_system()
            self.set_capture(False)
        return out
