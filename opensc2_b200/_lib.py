"""Loader of libopensc2_b200.so (the hand-written sm_100a kernels + C ABI).

There is no CPU fallback: a missing library or a missing CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

from ._abi import OSC2_K_COUNT, Osc2Model, Osc2StepInputs, Osc2Sweep, Osc2Topology, c_dp, c_ip

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libopensc2_b200.so")
_LIB = None


class Osc2Error(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"opensc2_b200 error {code}: {text}")
        self.code = code


class SingularMatrixError(ValueError):
    """Same condition the reference raises ValueError for (gredub/gbacsb, tsf:676-682, 763-786)."""


def load() -> C.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  opensc2_b200 has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    H = C.c_void_p
    L.osc2_last_error.restype = C.c_char_p
    L.osc2_version.restype = C.c_int
    L.osc2_device_count.restype = C.c_int
    L.osc2_create.argtypes = [C.POINTER(Osc2Topology), C.c_int, C.c_int, C.c_int, C.POINTER(H)]
    L.osc2_destroy.argtypes = [H]
    L.osc2_upload_state.argtypes = [H, c_dp, c_dp]
    L.osc2_upload_step_inputs.argtypes = [H, C.POINTER(Osc2StepInputs)]
    L.osc2_step.argtypes = [H, C.c_double, C.c_int]
    L.osc2_set_model.argtypes = [H, C.POINTER(Osc2Model)]
    L.osc2_set_fluid_table.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, c_dp]
    L.osc2_upload_sweep.argtypes = [H, C.POINTER(Osc2Sweep)]
    L.osc2_eval_operating_conditions.argtypes = [H, C.c_double]
    L.osc2_advance.argtypes = [H, C.c_double, C.c_double, C.c_int]
    L.osc2_eval_tcs.argtypes = [H]
    L.osc2_download_tcs.argtypes = [H, c_dp, c_dp]
    L.osc2_download_margin.argtypes = [H, c_dp]
    L.osc2_eval_nodal_properties.argtypes = [H]
    L.osc2_download_nodal_properties.argtypes = [H, c_dp]
    L.osc2_download_step_inputs.argtypes = [H] + [c_dp] * 8
    L.osc2_synchronize.argtypes = [H]
    L.osc2_download_state.argtypes = [H, c_dp, c_dp, c_ip]
    L.osc2_download_diagnostics.argtypes = [H, c_dp, c_dp, c_dp, c_dp]
    L.osc2_set_capture.argtypes = [H, C.c_int]
    L.osc2_download_system.argtypes = [H, c_dp, c_dp]
    L.osc2_set_profiling.argtypes = [H, C.c_int]
    L.osc2_get_kernel_ms.argtypes = [H, C.POINTER(C.c_double * OSC2_K_COUNT), C.POINTER(C.c_longlong * OSC2_K_COUNT)]
    L.osc2_reset_kernel_ms.argtypes = [H]
    L.osc2_timer_start.argtypes = [H]
    L.osc2_timer_stop.argtypes = [H, C.POINTER(C.c_double)]
    L.osc2_selftest_division.argtypes = [C.c_int, C.c_longlong, C.c_ulonglong, C.POINTER(C.c_ulonglong)]
    L.osc2_selftest_division.restype = C.c_int
    L.osc2_device_bytes.argtypes = [H]
    L.osc2_device_bytes.restype = C.c_size_t
    for name in ("osc2_create", "osc2_destroy", "osc2_upload_state", "osc2_upload_step_inputs", "osc2_step",
                 "osc2_synchronize", "osc2_download_state", "osc2_download_diagnostics", "osc2_set_capture",
                 "osc2_download_system", "osc2_set_profiling", "osc2_get_kernel_ms", "osc2_reset_kernel_ms",
                 "osc2_timer_start", "osc2_timer_stop", "osc2_set_model", "osc2_set_fluid_table", "osc2_upload_sweep",
                 "osc2_eval_operating_conditions", "osc2_download_step_inputs", "osc2_advance", "osc2_eval_tcs",
                 "osc2_download_tcs", "osc2_download_margin", "osc2_eval_nodal_properties",
                 "osc2_download_nodal_properties"):
        getattr(L, name).restype = C.c_int
    _LIB = L
    return L


EXPORTED = (
    "osc2_last_error", "osc2_version", "osc2_device_count", "osc2_create", "osc2_destroy",
    "osc2_upload_state", "osc2_upload_step_inputs", "osc2_step", "osc2_synchronize",
    "osc2_download_state", "osc2_download_diagnostics", "osc2_set_capture", "osc2_download_system",
    "osc2_set_profiling", "osc2_get_kernel_ms", "osc2_reset_kernel_ms", "osc2_timer_start",
    "osc2_timer_stop", "osc2_device_bytes", "osc2_selftest_division", "osc2_set_model", "osc2_set_fluid_table",
    "osc2_upload_sweep", "osc2_eval_operating_conditions", "osc2_download_step_inputs", "osc2_advance",
    "osc2_eval_tcs", "osc2_download_tcs", "osc2_download_margin", "osc2_eval_nodal_properties",
    "osc2_download_nodal_properties",
)


def check(rc: int) -> None:
    if rc == 0:
        return
    text = load().osc2_last_error().decode("utf-8", "replace")
    if rc == -4:
        raise SingularMatrixError(text)
    if rc in (-1, -7):          # -7: scipy.optimize.bisect's ValueError in the reference's Tcs evaluation
        raise ValueError(text)
    raise Osc2Error(rc, text)
