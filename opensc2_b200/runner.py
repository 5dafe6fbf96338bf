"""Transient runner: deck -> mesh -> initial state -> time loop on the device -> reference-format output.

The batched counterpart of the reference's driver sequence
  Simulation.conductor_initialization  (/root/reference/source_code/simulation.py:199-310,
                                        Conductor.initialization conductor.py:2679-2944)
  Simulation.conductor_solution        (simulation.py:312-480): while t < TEND: get_time_step ->
                                        operating_conditions_em / _th -> build_heat_source -> step -> post-step
                                        density and mass flow
  Simulation.conductor_post_processing (simulation.py:485-...): save_properties -> Solution/*.tsv
for B variants of ONE deck (swept heat-slug power, operating current, field, inlet temperature), all of them
advanced by one `osc2_advance` per time step.  The state never leaves the GPU between steps; what crosses PCIe per
step is nothing (fixed dt) or the d EQTEIG values per conductor the adaptive time step needs.

Not covered (refused, not approximated): several thermally coupled conductors (the reference raises too,
simulation.py:281-303), the electric module (a non-zero current enters the thermal path only through Tcs, as in
SURVEY.md Appendix C), negative (file-driven) flags.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from .batch import BatchedStep
from .initial_state import InitialState, initial_state
from .mesh import build_mesh
from .model import SOLID_STRAND_MIXED, FluidTable, Model, Sweep, heat_window
from .problem import BC_TINL, BC_TOUT, N_BC, StepArrays, Topology


def step_geometry(deck, zcoord: np.ndarray, batch: int, fl_bc: np.ndarray) -> StepArrays:
    """Everything the device does not compute: dz, constant contact perimeters at the Gauss points
    (conductor.py:1634-1830 with straight components: perimeter = the coupling-sheet value), qsource = 0
    (simulation.py:244-249), boundary values.  Coefficient arrays are zero-size placeholders."""
    topo = deck.topology
    N, E = zcoord.shape[0], zcoord.shape[0] - 1
    M, S = topo.n_ch, topo.n_sol
    nff, nfs, nss, nes = len(topo.ff), len(topo.fs), len(topo.ss), len(topo.es)
    ones = np.ones((batch, 1, E))
    dz = np.broadcast_to(zcoord[1:] - zcoord[:-1], (batch, E)).copy()
    peri_ff = np.empty((batch, 2, nff, E))
    if nff:
        peri_ff[:] = deck.peri["ff"][None, :, :, None]
    z = np.zeros
    # the coefficient arrays K_P computes on the device (and the state, kept elsewhere) are one-conductor placeholders:
    # at 4096 x 10 001 nodes they would be 20 GB of zeros on the host
    arr = StepArrays(
        dz=dz, fl_gauss=z((1, 9, M, E)), fl_node=z((1, 4, M)), fl_bc=np.array(fl_bc, dtype=np.float64),
        sol_gauss=z((1, 3, S, E)), sol_q=z((1, 2, S, E, 2)), peri_ff=peri_ff, htc_ff=z((1, 2, nff, E)),
        peri_fs=(deck.peri["fs"][None, :, None] * ones) if nfs else z((batch, 0, E)), htc_fs=z((1, nfs, E)),
        peri_ss=(deck.peri["ss"][None, :, None] * ones) if nss else z((batch, 0, E)), htc_ss=z((1, nss, E)),
        peri_es=(deck.peri["es"][None, :, None] * ones) if nes else z((batch, 0, E)), htc_es=z((1, nes, 3, E)),
        qsource=z((batch, N, S)), sysvar=z((1, N * topo.ndf)), syslod=z((1, N * topo.ndf, 2)),
        t_env=deck.t_env, dt=float(deck.transient.get("STPMIN", 0.1)), num_step=1)
    if any(abs(c[0]) == 3 for c in deck.model.ss):
        arr.peri_ss_nodal = deck.peri["ss"][None, :, None] * np.ones((batch, 1, N))
    return arr


def operating_current_density(deck, i_op: np.ndarray) -> np.ndarray:
    """(B, S) jop = |I_op * op_current_fraction_sh| / cross_section["sc"] for the strands holding a superconductor
    (solid_component.py:135-170: fraction = CROSSECTION / total strand cross-section; strand_mixed_component.py:
    255-262 with ISTAB_NON_STAB = 1: sc = CROSSECTION / (1 + STAB_NON_STAB); strand_component.py:266-276)."""
    topo, model = deck.topology, deck.model
    i_op = np.atleast_1d(np.asarray(i_op, dtype=np.float64))
    strands = [l for l, s in enumerate(model.solids) if s.kind != 2]
    total_so = sum(topo.sol_area[l] for l in strands)
    jop = np.zeros((i_op.shape[0], topo.n_sol))
    for l in strands:
        s = model.solids[l]
        if s.kind == SOLID_STRAND_MIXED:
            a_sc = topo.sol_area[l] / (1.0 + s.weights[0])
            jop[:, l] = np.abs(i_op * (topo.sol_area[l] / total_so)) / a_sc
    return jop


def deck_sweep(deck, zcoord: np.ndarray, batch: int, q0: Optional[np.ndarray] = None,
               b_field: Optional[np.ndarray] = None, i_op: Optional[np.ndarray] = None,
               eps: Optional[np.ndarray] = None) -> Sweep:
    """`Sweep` of a batch of variants of one deck.  q0: (B,) W/m on every heated solid (default: the deck's Q0);
    b_field: (B, 2) BISS, BOSS on every strand (default: the deck's); i_op: (B,) A total operating current
    (default 0 -> Tcs = Tc); eps: (B,) strain of the strands."""
    topo, model = deck.topology, deck.model
    S = topo.n_sol
    sw_b, sw_eps, sw_q0, win = np.zeros((batch, S, 2)), np.zeros((batch, S)), np.zeros((batch, S)), np.zeros((batch, S, 4))
    for l, o in enumerate(deck.solid_ops):
        sw_b[:, l, 0], sw_b[:, l, 1] = o["BISS"], o["BOSS"]
        sw_eps[:, l] = o["EPS"]
        if model.solids[l].kind != 2:
            if b_field is not None:
                sw_b[:, l, :] = b_field
            if eps is not None:
                sw_eps[:, l] = eps
        if o["IQFUN"] == 1:
            lo, hi = heat_window(zcoord, o["XQBEG"], o["XQEND"])
            win[:, l, :] = (lo, hi, o["TQBEG"], o["TQEND"])
            sw_q0[:, l] = o["Q0"] if q0 is None else q0
        else:
            win[:, l, :] = (0, -1, o["TQBEG"], o["TQEND"])     # TQBEG still feeds the transient HTC (conductor.py:4926)
    has_sc = any(s.kind == SOLID_STRAND_MIXED for s in model.solids)
    jop = operating_current_density(deck, np.zeros(batch) if i_op is None else i_op) if has_sc else None
    return Sweep(b_field=sw_b, eps=sw_eps, q0=sw_q0, q_window=win, tcs=None, jop=jop)


@dataclass
class TransientResult:
    time: List[float]
    sysvar: np.ndarray                       # (B, N*d) final state
    status: np.ndarray                       # (B,)
    snapshots: Dict[int, np.ndarray] = field(default_factory=dict)   # step -> (B, N*d)
    margin_min: Optional[np.ndarray] = None  # (B, S) min over the run of min_x (Tcs - T)
    tcs_nodal: Optional[np.ndarray] = None   # (B, S, N)
    nodal: Optional[Dict[str, np.ndarray]] = None   # run(want_nodal=True): {column: (B, M, N)}, the device's nodal coolant pass


class TransientSetup:
    """Host side of a run (no device): mesh, initial state, boundary values, sweep, geometry."""

    def __init__(self, deck, tables: Sequence[FluidTable], batch: int = 1, *, n_elems: Optional[int] = None,
                 q0=None, b_field=None, i_op=None, eps=None, t_inlet=None, legacy_intial2: bool = False):
        self.deck, self.tables, self.batch = deck, list(tables), int(batch)
        topo = deck.topology
        self.zcoord = build_mesh(deck.length, deck.mesh, n_elems)
        N = self.zcoord.shape[0]
        # ---- initial state (host, once): per distinct inlet temperature
        fl_bc = np.broadcast_to(deck.fl_bc, (self.batch, N_BC, topo.n_ch)).copy()
        if t_inlet is not None:
            fl_bc[:, BC_TINL, :] = np.asarray(t_inlet, dtype=np.float64)[:, None]
            fl_bc[:, BC_TOUT, :] = np.asarray(t_inlet, dtype=np.float64)[:, None]
        sysvar = np.empty((self.batch, N * topo.ndf))
        cache: Dict[bytes, InitialState] = {}
        for b in range(self.batch):
            key = fl_bc[b].tobytes()
            if key not in cache:
                cache[key] = initial_state(deck, self.tables, self.zcoord, fl_bc[b], legacy_intial2=legacy_intial2)
            sysvar[b], fl_bc[b] = cache[key].sysvar, cache[key].fl_bc
        self.fl_bc, self.sysvar0 = fl_bc, sysvar
        self.sweep = deck_sweep(deck, self.zcoord, self.batch, q0=q0, b_field=b_field, i_op=i_op, eps=eps)
        self.geometry = step_geometry(deck, self.zcoord, self.batch, fl_bc)
        self.has_sc = self.sweep.jop is not None


class Transient(TransientSetup):
    """B variants of one deck on one GPU."""

    def __init__(self, deck, tables: Sequence[FluidTable], batch: int = 1, *, device: int = 0,
                 get_time_step: Optional[Callable] = None, **kw):
        super().__init__(deck, tables, batch, **kw)
        topo, N = deck.topology, self.zcoord.shape[0]
        self.bs = BatchedStep(topo, self.batch, N, device=device)
        self.bs.set_model(deck.model, self.tables)
        self.bs.upload_sweep(self.sweep)
        self.bs.upload_state(self.sysvar0, None)          # SYSLOD starts at zero on the device (osc2_create)
        self.bs.upload_geometry(self.geometry)
        if self.has_sc:
            self.bs.eval_tcs()                     # once, at step 0 (conductor.py:4615-4619)
        self.time = [0.0]
        self.num_step = 0
        self._adaptive = int(deck.transient.get("IADAPTIME", 0)) != 0
        self._get_time_step = get_time_step

    def close(self):
        self.bs.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def time_step(self) -> float:
        """get_time_step (transient_solution_functions.py:38-184); IADAPTIME = 0: min(STPMIN, TEND - t)."""
        tr = self.deck.transient
        if not self._adaptive:
            return float(min(tr["STPMIN"], tr["TEND"] - self.time[-1])) if "TEND" in tr else float(tr["STPMIN"])
        if self._get_time_step is None:
            raise ValueError("IADAPTIME != 0 needs a get_time_step callback (opensc2_b200.time_step.get_time_step "
                             "on a per-conductor view): the batch advances with one common dt")
        return float(self._get_time_step(self))

    def run(self, t_end: Optional[float] = None, max_steps: Optional[int] = None, snapshot_every: int = 0,
            track_margin: bool = False, on_step: Optional[Callable] = None, want_nodal: bool = False) -> TransientResult:
        """want_nodal: also run the device's nodal coolant pass on the final state (post-step density, mass flow rate,
        the property columns of save_properties; 14 M N doubles per conductor come back to the host)."""
        tr = self.deck.transient
        t_end = float(tr["TEND"]) if t_end is None else float(t_end)
        stpmin = float(tr["STPMIN"])
        snaps: Dict[int, np.ndarray] = {}
        margin = None
        steps = 0
        # simulation.py:341-345: while t < TEND - 1e-5 * STPMIN
        while self.time[-1] < t_end - 1e-5 * stpmin and (max_steps is None or steps < max_steps):
            dt = min(self.time_step(), t_end - self.time[-1]) if not self._adaptive else self.time_step()
            self.num_step += 1
            t_new = self.time[-1] + dt
            self.time.append(t_new)
            self.bs.advance(t_new, dt, self.num_step)
            steps += 1
            if track_margin and self.has_sc:
                m = self.bs.download_margin()
                margin = m if margin is None else np.minimum(margin, m)
            if snapshot_every and self.num_step % snapshot_every == 0:
                snaps[self.num_step] = self.bs.download_state(want_syslod=False)[0]
            if on_step is not None:
                on_step(self)
        self.bs.synchronize()                      # raises like the reference on a singular system
        sv, _, st = self.bs.download_state(want_syslod=False)
        tcs_nodal = self.bs.download_tcs()[1] if self.has_sc else None
        # post-step density, mass flow rate and the nodal property columns of the final state, on the device
        # (coolant.py:226-263; what save_properties prints)
        nodal = self.bs.nodal_properties() if (want_nodal and self.deck.topology.n_ch > 0) else None
        return TransientResult(time=list(self.time), sysvar=sv, status=st, snapshots=snaps, margin_min=margin,
                               tcs_nodal=tcs_nodal, nodal=nodal)

    def save_solution(self, folder: str, result: TransientResult, conductors: Optional[Sequence[int]] = None):
        """Solution/*.tsv per conductor, the reference's save_properties layout (output.py)."""
        import os

        from .output import save_properties

        for b in (range(self.batch) if conductors is None else conductors):
            save_properties(os.path.join(folder, f"conductor_{b}") if self.batch > 1 else folder, self.deck.topology,
                            self.deck.model, self.tables, self.zcoord, result.sysvar[b], self.sweep.b_field[b],
                            None if result.tcs_nodal is None else result.tcs_nodal[b],
                            chan_ids=self.deck.chan_ids, solid_ids=self.deck.solid_ids,
                            nodal=None if result.nodal is None else {k: v[b] for k, v in result.nodal.items()})
