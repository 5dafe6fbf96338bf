"""`.tsv` writers on the far side of the path: drop-in for the reference's `save_properties`
(/root/reference/source_code/utility_functions/output.py:8-92) fed from batched device state.

One file per component, one row per node, `%.18e`, the column layout of HEAD:
  channels  zcoord, pressure, temperature, total_density, velocity, the ten coolant properties in the
            order of the alias table (simulation.py:89-100), Reynolds, Gruneisen, mass_flow_rate,
            friction_factor
  strands   zcoord, temperature, B_field [, T_cur_sharing for components holding a superconductor]
  jackets   zcoord, temperature
Host-side post-processing (numpy): nothing here is on the time-step path.  Nodal coolant properties use
the same bilinear table lookup as the device (`model.table_lookup`)."""
from __future__ import annotations

import os
from typing import Optional, Sequence

import numpy as np

from .model import (FP_BETA, FP_CP, FP_CV, FP_H, FP_K, FP_KAPPA, FP_MU, FP_PRANDTL, FP_RHO, FP_SOUND, SOLID_JACKET,
                    SOLID_STACK, SOLID_STRAND_MIXED, FluidTable, Model, table_lookup)
from .correlations import channel_friction_factor
from .problem import Topology

CHAN_COLUMNS = (
    ("zcoord", "(m)"), ("pressure", "(Pa)"), ("temperature", "(K)"), ("total_density", "(kg/m^3)"),
    ("velocity", "(m/s)"), ("isobaric_expansion_coefficient", "(1/K)"), ("isothermal_compressibility", "(1/Pa)"),
    ("Prandtl", "(~)"), ("total_dynamic_viscosity", "(Pa*s)"), ("total_enthalpy", "(J/kg)"),
    ("total_isobaric_specific_heat", "(J/kg/K)"), ("total_isochoric_specific_heat", "(J/kg/K)"),
    ("total_speed_of_sound", "(m/s)"), ("total_thermal_conductivity", "(W/m/K)"), ("Reynolds", "(~)"),
    ("Gruneisen", "(~)"), ("mass_flow_rate", "(kg/s)"), ("friction_factor", "(~)"),
)
_ALIAS_ORDER = (FP_BETA, FP_KAPPA, FP_PRANDTL, FP_MU, FP_H, FP_CP, FP_CV, FP_SOUND, FP_K)


def nodal_friction_factor(model: Model, topo: Topology, j: int, reynolds: np.ndarray) -> np.ndarray:
    """`Channel.eval_friction_factor(reynolds, nodal=True)` (channel.py:1119-1138): every IFRICTION flag the
    device accepts (`correlations.friction_factor`)."""
    return channel_friction_factor(model, topo, j, reynolds)


NODAL_COLUMNS = ("total_density", "isobaric_expansion_coefficient", "isothermal_compressibility", "Prandtl",
                 "total_dynamic_viscosity", "total_enthalpy", "total_isobaric_specific_heat", "total_isochoric_specific_heat",
                 "total_speed_of_sound", "total_thermal_conductivity", "Reynolds", "Gruneisen", "mass_flow_rate",
                 "friction_factor")         # BatchedStep.nodal_properties(): the device's nodal coolant pass


def channel_table(topo: Topology, model: Model, tab: FluidTable, j: int, zcoord, v, p, t, nodal=None) -> np.ndarray:
    """(N, 18) array of one channel's file.  `nodal`: {column: (M, N)} of this conductor from the device's nodal pass
    (osc2_eval_nodal_properties); without it the columns are evaluated here with the same table lookup."""
    if nodal is not None:
        return np.stack([zcoord, p, t, nodal["total_density"][j], v] + [nodal[k][j] for k in NODAL_COLUMNS[1:]], axis=1)
    rho = table_lookup(tab, FP_RHO, p, t)
    cols = [zcoord, p, t, rho, v] + [table_lookup(tab, q, p, t) for q in _ALIAS_ORDER]
    mu, beta, kappa, cv = cols[8], cols[5], cols[6], cols[11]
    reynolds = np.abs(topo.ch_dh[j] * v * rho / mu)                   # coolant.py:165-180
    grun = beta / kappa / cv / rho
    mdot = topo.ch_area[j] * v * rho                                   # coolant.py:247-263
    cols += [reynolds, grun, mdot, nodal_friction_factor(model, topo, j, reynolds)]
    return np.stack(cols, axis=1)


def save_properties(f_path: str, topo: Topology, model: Model, tables: Sequence[FluidTable], zcoord: np.ndarray,
                    sysvar: np.ndarray, b_field: np.ndarray, t_cur_sharing: Optional[np.ndarray] = None,
                    chan_ids: Optional[Sequence[str]] = None, solid_ids: Optional[Sequence[str]] = None,
                    nodal=None) -> None:
    """Write `<identifier>.tsv` for every component of ONE conductor.

    sysvar: (N*d,) state; b_field: (S, 2) BISS, BOSS; t_cur_sharing: (S, N) nodal current sharing
    temperature of the superconducting strands (rows of other solids are ignored); nodal: this conductor's
    {column: (M, N)} from the device's nodal coolant pass (optional, see channel_table)."""
    M, S, d = topo.n_ch, topo.n_sol, topo.ndf
    N = zcoord.shape[0]
    x = np.asarray(sysvar, dtype=np.float64).reshape(N, d)
    chan_ids = chan_ids or [f"CHAN_{j + 1}" for j in range(M)]
    if solid_ids is None:
        names = {SOLID_STRAND_MIXED: "STR_MIX", SOLID_JACKET: "Z_JACKET", SOLID_STACK: "STACK"}
        count, solid_ids = {}, []
        for s in model.solids:
            nm = names.get(s.kind, "STR_STAB")
            count[nm] = count.get(nm, 0) + 1
            solid_ids.append(f"{nm}_{count[nm]}")
    os.makedirs(f_path, exist_ok=True)
    header_chan = "\t".join(f"{n} {u}" for n, u in CHAN_COLUMNS)
    for j in range(M):
        a = channel_table(topo, model, tables[model.ch_fluid[j]], j, zcoord, x[:, j].copy(), x[:, M + j].copy(),
                          x[:, 2 * M + j].copy(), nodal)
        with open(os.path.join(f_path, f"{chan_ids[j]}.tsv"), "w") as w:
            np.savetxt(w, a, delimiter="\t", header=header_chan, comments="")
    for l, s in enumerate(model.solids):
        temp = x[:, 3 * M + l]
        if s.kind == SOLID_JACKET:
            a, header = np.stack([zcoord, temp], axis=1), "zcoord (m)\ttemperature (K)"
        else:
            bf = np.linspace(b_field[l, 0], b_field[l, 1], N)             # solid_component.py:383-388
            if s.kind in (SOLID_STRAND_MIXED, SOLID_STACK):     # every strand that is not a stabiliser
                tcs = np.zeros(N) if t_cur_sharing is None else t_cur_sharing[l]
                a = np.stack([zcoord, temp, bf, tcs], axis=1)
                header = "zcoord (m)\ttemperature (K)\tB_field (T)\tT_cur_sharing (K)"
            else:
                a, header = np.stack([zcoord, temp, bf], axis=1), "zcoord (m)\ttemperature (K)\tB_field (T)"
        with open(os.path.join(f_path, f"{solid_ids[l]}.tsv"), "w") as w:
            np.savetxt(w, a, delimiter="\t", header=header, comments="")


def save_properties_batch(root: str, topo: Topology, model: Model, tables, zcoord, sysvar, sweep,
                          conductors: Optional[Sequence[int]] = None, t_cur_sharing=None) -> None:
    """`save_properties` for slots of a batch: `<root>/conductor_<b>/...`; sysvar: (B, N*d)."""
    for b in (range(sysvar.shape[0]) if conductors is None else conductors):
        save_properties(os.path.join(root, f"conductor_{b}"), topo, model, tables, zcoord, sysvar[b], sweep.b_field[b],
                        None if t_cur_sharing is None else t_cur_sharing[b])


_SD_CHAN_HEADER = ("zcoord (m)\tvelocity (m/s)\tpressure (Pa)\ttemperature (K)\ttotal_density (kg/m^3)\t"
                   "friction_factor (~)")
_SD_GAUSS_HEADER = "zcoord_gauss (m)\tcurrent_along (A)\tdelta_voltage_along (V)\tP_along (W/m)"
_SD_TABLES = (("heat_rad_inner", "Heat_rad_inner"), ("heat_exch_env", "Heat_exch_env"), ("htc_ch_ch_open", "HTC_ch_ch_o"),
              ("htc_ch_ch_close", "HTC_ch_ch_c"), ("htc_ch_sol", "HTC_ch_sol"), ("htc_sol_sol_cond", "HTC_sol_sol_cond"),
              ("htc_sol_sol_rad", "HTC_sol_sol_rad"))


def save_simulation_space(f_path: str, topo: Topology, model: Model, tables: Sequence[FluidTable], zcoord: np.ndarray,
                          sysvar: np.ndarray, cond_num_step: int, time: float, first_store: bool, *,
                          zcoord_gauss: Optional[np.ndarray] = None, j_critical: Optional[np.ndarray] = None,
                          t_cur_sharing: Optional[np.ndarray] = None, tcs_evaluation: bool = False,
                          electric: Optional[np.ndarray] = None, nodal_tables: Optional[dict] = None,
                          chan_ids: Optional[Sequence[str]] = None, solid_ids: Optional[Sequence[str]] = None) -> None:
    """Spatial distributions of ONE conductor at one stored time: drop-in for the reference's
    `save_simulation_space` (utility_functions/output.py:98-347) — same file names, headers, formats.

      <channel>_(<step>)_sd.tsv        zcoord, velocity, pressure, temperature, total_density, friction_factor
      <strand>_(<step>)_sd.tsv         zcoord, temperature [, current sharing temperature], critical current density
                                       (stabilisers: zcoord, temperature)
      <jacket>_(<step>)_sd.tsv         zcoord, temperature
      <solid>_(<step>)_gauss_sd.tsv    zcoord_gauss, current, voltage drop, Joule power per length (electric module
                                       arrays; zeros when `electric` is None, as with the electric module off)
      <table>_(<step>)_sd.tsv          optional nodal tables (`nodal_tables`: heat_rad_inner, heat_exch_env,
                                       htc_ch_ch_open, htc_ch_ch_close, htc_ch_sol, htc_sol_sol_cond, htc_sol_sol_rad;
                                       each {column name: (N,) array}), written with pandas like the reference
      Time_sd_actual.tsv               created at the first store, appended afterwards

    sysvar: (N*d,); j_critical, t_cur_sharing: (S, N) (rows of non-superconducting solids ignored);
    electric: (S, E, 3) current_along, delta_voltage_along, linear_power_el_resistance[:, 0]."""
    import pandas as pd

    M, d = topo.n_ch, topo.ndf
    N = zcoord.shape[0]
    x = np.asarray(sysvar, dtype=np.float64).reshape(N, d)
    zg = (zcoord[:-1] + zcoord[1:]) / 2.0 if zcoord_gauss is None else zcoord_gauss
    chan_ids = chan_ids or [f"CHAN_{j + 1}" for j in range(M)]
    if solid_ids is None:
        names = {SOLID_STRAND_MIXED: "STR_MIX", SOLID_JACKET: "Z_JACKET", SOLID_STACK: "STACK"}
        count, solid_ids = {}, []
        for s in model.solids:
            nm = names.get(s.kind, "STR_STAB")
            count[nm] = count.get(nm, 0) + 1
            solid_ids.append(f"{nm}_{count[nm]}")
    os.makedirs(f_path, exist_ok=True)

    def write(name, a, header):
        with open(os.path.join(f_path, name), "w") as w:
            np.savetxt(w, a, delimiter="\t", header=header, comments="")

    for j in range(M):
        full = channel_table(topo, model, tables[model.ch_fluid[j]], j, zcoord, x[:, j].copy(), x[:, M + j].copy(),
                             x[:, 2 * M + j].copy())
        # channel_table columns: zcoord, pressure, temperature, total_density, velocity, ..., friction_factor
        write(f"{chan_ids[j]}_({cond_num_step})_sd.tsv", full[:, [0, 4, 1, 2, 3, -1]], _SD_CHAN_HEADER)
    for l, s in enumerate(model.solids):
        temp = x[:, 3 * M + l]
        if s.kind == SOLID_JACKET:
            a, header = np.stack([zcoord, temp], axis=1), "zcoord (m)\ttemperature (K)"
        elif s.kind == SOLID_STRAND_MIXED:
            jc = np.zeros(N) if j_critical is None else j_critical[l]
            if tcs_evaluation:
                tcs = np.zeros(N) if t_cur_sharing is None else t_cur_sharing[l]
                a = np.stack([zcoord, temp, tcs, jc], axis=1)
                header = ("zcoord (m)\ttemperature (K)\tcurrent_sharing_temperature (K)\t"
                          "critical_current_density (A/m^2)")
            else:
                a = np.stack([zcoord, temp, jc], axis=1)
                header = "zcoord (m)\ttemperature (K)\tcritical_current_density (A/m^2)"
        else:
            a, header = np.stack([zcoord, temp], axis=1), "zcoord (m)\ttemperature (K)"
        write(f"{solid_ids[l]}_({cond_num_step})_sd.tsv", a, header)
    for l in range(topo.n_sol):
        el = np.zeros((N - 1, 3)) if electric is None else electric[l]
        write(f"{solid_ids[l]}_({cond_num_step})_gauss_sd.tsv", np.concatenate([zg[:, None], el], axis=1), _SD_GAUSS_HEADER)
    for key, stem in _SD_TABLES:
        tab = (nodal_tables or {}).get(key)
        if tab:
            pd.DataFrame.from_dict(tab, dtype=float).to_csv(os.path.join(f_path, f"{stem}_({cond_num_step})_sd.tsv"),
                                                            sep="\t", index=False, header=True)
    with open(os.path.join(f_path, "Time_sd_actual.tsv"), "w" if first_store else "a") as w:
        if first_store:
            np.savetxt(w, np.array([time]), header="time (s)", comments="", delimiter="\t")
        else:
            np.savetxt(w, np.array([time]), comments="", delimiter="\t")


_IO_KEYS = ("velocity", "pressure", "temperature", "total_density", "mass_flow_rate")
_IO_UNITS = {"velocity": "(m/s)", "pressure": "(Pa)", "temperature": "(K)", "total_density": "(kg/m^3)",
             "mass_flow_rate": "(kg/s)"}
_GAUSS_KEYS = ("current_along", "delta_voltage_along", "linear_power_el_resistance", "delta_voltage_along_sum")


class TimeEvolutionWriter:
    """Time evolutions of ONE conductor at the user's axial coordinates: drop-in for the reference's
    `save_simulation_time` (utility_functions/output.py:647-977) with its buffering — header-only files at the
    first call, rows appended every `chunk_size` calls (`Conductor.CHUNCK_SIZE`, conductor.py:97) and when
    TEND is reached, `Time.tsv` at TEND.

      <channel>_{velocity,pressure,temperature,total_density,friction_factor}_te.tsv   value at each coordinate
      <channel>_inlet_outlet_te.tsv      v, p, T, rho, mass flow rate at the inlet and at the outlet node
      <solid>_<key>_te.tsv               temperature (+ B_field, T_cur_sharing, J_critical for strands: mixed
                                         strands all four, stabilisers the first two) and the Gauss-point arrays of
                                         the electric module (current_along, delta_voltage_along,
                                         linear_power_el_resistance[:, 0], delta_voltage_along_sum; jackets: first three)
    The coordinate -> node rule is the reference's: the last node with z <= round(coordinate, n_digit_z); the
    first Gauss coordinate always maps to element 0 (output.py:661-678)."""

    def __init__(self, f_path: str, topo: Topology, model: Model, tables: Sequence[FluidTable], time_save: np.ndarray,
                 tend: float, *, n_digit_z: int = 3, chunk_size: int = 100, ch_forward: Optional[Sequence[bool]] = None,
                 chan_ids: Optional[Sequence[str]] = None, solid_ids: Optional[Sequence[str]] = None):
        self.f_path, self.topo, self.model, self.tables = f_path, topo, model, tables
        self.time_save, self.tend, self.n_digit_z, self.chunk = np.asarray(time_save, dtype=float), tend, n_digit_z, chunk_size
        self.ch_forward = list(ch_forward) if ch_forward is not None else [bool(f) for f in topo.ch_forward]
        self.chan_ids = list(chan_ids) if chan_ids else [f"CHAN_{j + 1}" for j in range(topo.n_ch)]
        if solid_ids is None:
            names = {SOLID_STRAND_MIXED: "STR_MIX", SOLID_JACKET: "Z_JACKET", SOLID_STACK: "STACK"}
            count, solid_ids = {}, []
            for s in model.solids:
                nm = names.get(s.kind, "STR_STAB")
                count[nm] = count.get(nm, 0) + 1
                solid_ids.append(f"{nm}_{count[nm]}")
        self.solid_ids = list(solid_ids)
        self.cond_time: list = []
        self.buf: dict = {}          # file name -> {column: list}
        self.started = False

    def _solid_keys(self, l):
        kind = self.model.solids[l].kind
        if kind == SOLID_JACKET:
            return ("temperature",), _GAUSS_KEYS[:3]
        if kind == SOLID_STRAND_MIXED:
            return ("temperature", "B_field", "T_cur_sharing", "J_critical"), _GAUSS_KEYS
        return ("temperature", "B_field"), _GAUSS_KEYS

    def _path(self, name):
        return os.path.join(self.f_path, name)

    def _flush(self, name, at_tend):
        import pandas as pd

        val = self.buf[name]
        if len(val["time (s)"]) == self.chunk or at_tend:
            pd.DataFrame(val, columns=list(val.keys()), dtype=float).to_csv(self._path(name), sep="\t", mode="a",
                                                                            chunksize=self.chunk, index=False, header=False)
            if len(val["time (s)"]) == self.chunk:
                self.buf[name] = {k: [] for k in val}

    def store(self, time: float, zcoord: np.ndarray, sysvar: np.ndarray, *, zcoord_gauss: Optional[np.ndarray] = None,
              solid_nodal: Optional[dict] = None, solid_gauss: Optional[dict] = None) -> None:
        """One call per time step.  sysvar: (N*d,); solid_nodal: {key: (S, N)} for B_field, T_cur_sharing,
        J_critical; solid_gauss: {key: (S, E)} for the electric-module arrays (missing ones are zeros)."""
        import pandas as pd

        topo, model = self.topo, self.model
        M, S, d = topo.n_ch, topo.n_sol, topo.ndf
        N = zcoord.shape[0]
        x = np.asarray(sysvar, dtype=np.float64).reshape(N, d)
        zg = (zcoord[:-1] + zcoord[1:]) / 2.0 if zcoord_gauss is None else zcoord_gauss
        ts = self.time_save
        ind = {f"zcoord = {ts[i]} (m)": int(np.max(np.nonzero(zcoord <= round(ts[i], self.n_digit_z)))) for i in range(ts.size)}
        ind_g = {f"zcoord_g = {ts[0]} (m)": 0}
        ind_g.update({f"zcoord_g = {ts[i]} (m)": int(np.max(np.nonzero(zg <= round(ts[i], self.n_digit_z)))) for i in range(1, ts.size)})
        io_cols = ["time (s)"] + [f"{k}_{side} {_IO_UNITS[k]}" for side in ("inl", "out") for k in _IO_KEYS]
        if not self.started:
            os.makedirs(self.f_path, exist_ok=True)

            def create(name, columns):
                pd.DataFrame(columns=columns).to_csv(self._path(name), sep="\t", index=False, header=True)
                self.buf[name] = {c: [] for c in columns}

            for j in range(M):
                for key in ("velocity", "pressure", "temperature", "total_density", "friction_factor"):
                    create(f"{self.chan_ids[j]}_{key}_te.tsv", ["time (s)"] + list(ind))
                create(f"{self.chan_ids[j]}_inlet_outlet_te.tsv", io_cols)
            for l in range(S):
                nodal, gauss = self._solid_keys(l)
                for key in nodal:
                    create(f"{self.solid_ids[l]}_{key}_te.tsv", ["time (s)"] + list(ind))
                for key in gauss:
                    create(f"{self.solid_ids[l]}_{key}_te.tsv", ["time (s)"] + list(ind_g))
            self.started = True
        self.cond_time.append(time)
        at_tend = abs(time - self.tend) / self.tend <= 1e-6

        def push(name, prop, index):
            val = self.buf[name]
            val["time (s)"].append(time)
            for key, i in index.items():
                val[key].append(prop[i])
            self._flush(name, at_tend)

        for j in range(M):
            full = channel_table(topo, model, self.tables[model.ch_fluid[j]], j, zcoord, x[:, j].copy(), x[:, M + j].copy(),
                                 x[:, 2 * M + j].copy())
            cols = {"velocity": full[:, 4], "pressure": full[:, 1], "temperature": full[:, 2], "total_density": full[:, 3],
                    "mass_flow_rate": full[:, 16], "friction_factor": full[:, 17]}
            for key in ("velocity", "pressure", "temperature", "total_density", "friction_factor"):
                push(f"{self.chan_ids[j]}_{key}_te.tsv", cols[key], ind)
            i_in, i_out = (0, -1) if self.ch_forward[j] else (-1, 0)
            name = f"{self.chan_ids[j]}_inlet_outlet_te.tsv"
            val = self.buf[name]
            val["time (s)"].append(time)
            for side, i in (("inl", i_in), ("out", i_out)):
                for k in _IO_KEYS:
                    val[f"{k}_{side} {_IO_UNITS[k]}"].append(cols[k][i])
            self._flush(name, at_tend)
        for l in range(S):
            nodal, gauss = self._solid_keys(l)
            for key in nodal:
                prop = x[:, 3 * M + l] if key == "temperature" else (
                    np.zeros(N) if not solid_nodal or key not in solid_nodal else solid_nodal[key][l])
                push(f"{self.solid_ids[l]}_{key}_te.tsv", prop, ind)
            for key in gauss:
                prop = np.zeros(N - 1) if not solid_gauss or key not in solid_gauss else solid_gauss[key][l]
                push(f"{self.solid_ids[l]}_{key}_te.tsv", prop, ind_g)
        if at_tend:
            pd.Series(self.cond_time, name="time (s)", dtype=float).to_csv(self._path("Time.tsv"), sep="\t", header=True,
                                                                           index=False)
