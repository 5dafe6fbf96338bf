#!/usr/bin/env python
"""Benchmark of the batched OPENSC2 time-step path (driver contract in the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload config3|config4]

Metric: node-timesteps/s = conductors x nodes x steps / device time, whole job.

`--workload config3` (default; BASELINE.json configs[3]): 4096 CASE_1 variants per GPU on a 10 001-node mesh, built
by the SAME pipeline as a real run (`opensc2_b200.runner`: deck image of TDD_examples/CASE_1 -> mesh -> gen_flow ->
initial state), swept in heat-slug power, operating current (-> Tcs by the device bisection), background field and
inlet temperature.  One "step" = one full transient time step: operating conditions at the Gauss points (coolant
tables, correlations, material fits, heat source, temperature margin; K_P) + step() (assembly, boundary rows, scaling,
banded LU, substitution, norms; K_G, K_F, K_B, K_N).  N > 1: every rank owns its own 4096 conductors (weak scaling, no
collective on the data path).

`--workload config4` (BASELINE.json configs[4]): the temperature-margin sweep of 131 072 variants -- 65 536 of CASE_1
(201 nodes) + 65 536 of CASE_2 (NODOFS 64, 151 nodes) -- sharded over the N ranks (STRONG scaling), CASE_2 in chunks
that fit HBM, final all_gather of the solution fields.

`--impl reference` times the CPU oracle port of the same work (oracle/props_oracle.c + oracle/step_oracle.c) on all
host threads, rank 0 only: the reference itself is Python and cannot travel to the GPU box.
"""
from __future__ import annotations

import argparse
import dataclasses
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from opensc2_b200.deck import deck_from_dict  # noqa: E402
from opensc2_b200.model import synthetic_helium_table  # noqa: E402
from opensc2_b200.runner import Transient, TransientSetup  # noqa: E402

BATCH = int(os.environ.get("OSC2_BENCH_BATCH", 4096))
NODES = int(os.environ.get("OSC2_BENCH_NODES", 10001))
N_INLET = 256        # distinct inlet temperatures (one gen_flow + initial state each); every other swept scalar is per conductor
T_START = 9.5        # s: the timed steps run through the start of the heat slug (TQBEG = 10 s)
DECKS = os.path.join(ROOT, "tests", "golden", "decks_summary.json")


def deck_image(case):
    with open(DECKS) as f:
        return deck_from_dict(json.load(f)[case])


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def b_alg(d: int) -> int:
    """algorithmic bytes per node-timestep, two-kernel model."""
    return 8 * (3 * d * (4 * d - 1) + 6 * d)


def touched_smat_entries(topo) -> int:
    """Entries of the d x d source Jacobian a statement of K_G can write (osc2_build_smap, step_launch.h)."""
    M = topo.n_ch
    e = set()
    for j in range(M):
        e |= {(j, j), (M + j, j), (2 * M + j, j)}
    for (a, b) in topo.ff:
        for c1, c2 in ((a, b), (b, a)):
            v1, p1, t1, p2, t2 = c1, M + c1, 2 * M + c1, M + c2, 2 * M + c2
            e |= {(v1, p1), (v1, p2), (p1, p1), (p1, p2), (p1, t1), (p1, t2), (t1, p1), (t1, p2), (t1, t1), (t1, t2)}
    for (j, l) in topo.fs:
        ip, it, is_ = M + j, 2 * M + j, 3 * M + l
        e |= {(ip, it), (ip, is_), (it, it), (it, is_), (is_, is_), (is_, it)}
    for (a, c) in topo.ss:
        a, c = 3 * M + a, 3 * M + c
        e |= {(a, a), (a, c), (c, c), (c, a)}
    for l in topo.es:
        e.add((3 * M + l, 3 * M + l))
    return len(e)


def kernel_alg_bytes(topo):
    """Per-unit algorithmic bytes of each kernel."""
    d, M, S = topo.ndf, topo.n_ch, topo.n_sol
    nff, nfs, nss, nes = len(topo.ff), len(topo.fs), len(topo.ss), len(topo.es)
    rec = d * (d + 9)
    # record fields K_G writes every step: the ones that are +0 for every element of the topology (untouched SMAT
    # entries, load halves of the fluid rows, advection entries of the solid rows) were written once, at creation
    rec_written = rec - ((d * d - touched_smat_entries(topo)) + 4 * 3 * M + 2 * S)
    coef = 1 + 9 * M + 3 * S + 4 * S + 4 * nff + 2 * nfs + 2 * nss + 4 * nes     # per element, K_G inputs
    computed = 9 * M + 3 * S + 4 * S + 2 * nff + nfs + nss + 3 * nes              # per element, K_P outputs
    return {
        "props": 8 * (d + S + computed),                    # per element: state, Tcs in; coefficients out
        "gauss": 8 * (coef + S + rec_written + 3 * nff),    # per element: coefficients, qsource in; row records, K1..K3 out
        "forward": 8 * (rec + d + 4 * d + d * (2 * d + 2)),  # per node: record, x_old, SYSLOD in/out, factors out
        "backward": 8 * (d * (2 * d + 2) + d),              # per node: factors in, x out
        "norms": 8 * d,                                      # per node: x in
    }


def csrc_hash():
    h = hashlib.sha1()
    d = os.path.join(ROOT, "opensc2_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(d, f), "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()[:12]


def measured_traffic(kernel_family):
    """DRAM bytes per unit of the dominant kernel from the last `ncu --set full` capture (tools/ncu_traffic.py writes
    profiles/ncu_traffic.json); None when the kernels changed since that capture."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, "no ncu capture on file"
    with open(p) as f:
        j = json.load(f)
    if j.get("csrc_sha1") != csrc_hash():
        return None, f"ncu capture {j.get('report')} predates the current kernels (csrc hash differs): not quoted"
    e = j["kernels"].get(kernel_family)
    if not e:
        return None, "kernel not in the capture"
    return float(e["dram_bytes_per_unit"]), f"ncu --set full, dram__bytes_read+write of {e['name']} at this size ({j.get('report')})"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# configs[3]
# ---------------------------------------------------------------------------------------------------------
def sweep_values(batch, seed):
    """Per-conductor swept scalars of SURVEY.md 8(d): slug power, operating current, background field, inlet T.
    The current stays below Jc(3.5 K) at the largest field of the sweep (the reference's Tcs bisection raises above)."""
    rng = np.random.default_rng(seed)
    b0 = rng.uniform(0.0, 11.0, batch)
    return dict(q0=rng.uniform(50.0, 1000.0, batch), i_op=rng.uniform(0.0, 60.0e3, batch),
                b_field=np.stack([b0, b0 * rng.uniform(0.9, 1.1, batch)], axis=1),
                t_inlet=rng.uniform(4.3, 4.7, N_INLET)[rng.integers(0, N_INLET, batch)])


def config3_setup(batch, nodes, seed, device=None):
    deck, tab = deck_image("CASE_1_ITER_like_LTS"), synthetic_helium_table()
    kw = sweep_values(batch, seed)
    cls = TransientSetup if device is None else Transient
    extra = {} if device is None else {"device": device}
    tr = cls(deck, [tab], batch, n_elems=nodes - 1, **kw, **extra)
    return deck, tab, tr


def workload_config(topo):
    return {"workload": f"BASELINE configs[3]: {BATCH} CASE_1 variants per GPU (d={topo.ndf}: 2 He channels + Cu/Nb3Sn strand + "
                        f"SS/glass-epoxy jacket, 5 interfaces), {NODES}-node mesh, Backward Euler dt=0.1 s; every conductor has its "
                        f"own slug power, operating current (Tcs by the device bisection), field; {N_INLET} distinct inlet "
                        "temperatures (one gen_flow + initial state each); deck -> mesh -> gen_flow -> initial state through "
                        "opensc2_b200.runner; one step = operating conditions at the Gauss points (tables, fits, heat source, "
                        "temperature margin) + step()",
            "batch_per_gpu": BATCH, "nodes": NODES, "ndf": topo.ndf,
            "coolant_table": "synthetic helium stand-in (CoolProp absent; DESIGN.md 3)",
            "l2": "state, coefficients and factors (>100 GB per step) exceed the 126 MB L2, no flush needed"}


def oracle_steps(tr, tcs, n_cond, n_steps, threads, t_start):
    """The oracle's own loop on the first `n_cond` conductors of the set-up: (node-timesteps/s, seconds, final state)."""
    from oracle import oracle as O

    O.lib()
    topo, model = tr.deck.topology, tr.deck.model
    N = tr.zcoord.shape[0]
    arr = tr.geometry.copy()
    for f in dataclasses.fields(arr):
        v = getattr(arr, f.name)
        if isinstance(v, np.ndarray) and v.ndim and v.shape[0] == tr.batch:
            setattr(arr, f.name, np.ascontiguousarray(v[:n_cond]))
    arr.sysvar, arr.syslod = tr.sysvar0[:n_cond].copy(), np.zeros((n_cond, N * topo.ndf, 2))
    sw = type(tr.sweep)(**{f.name: (getattr(tr.sweep, f.name)[:n_cond] if getattr(tr.sweep, f.name) is not None else None)
                            for f in dataclasses.fields(tr.sweep)})
    sw.tcs = None if tcs is None else np.ascontiguousarray(tcs[:n_cond])
    dt = float(tr.deck.transient["STPMIN"])
    arr.dt = dt
    t, t0 = t_start, time.perf_counter()
    for k in range(1, n_steps + 1):
        t += dt
        ins = O.operating_conditions(topo, model, tr.tables, sw, arr.sysvar, N, t, n_threads=threads)
        for key, v in ins.items():
            setattr(arr, key, v)
        arr.num_step = k
        out = O.step(topo, arr, n_threads=threads)
        arr.sysvar, arr.syslod = out["sysvar"], out["syslod"]
    sec = time.perf_counter() - t0
    return n_cond * N * n_steps / sec, sec, arr.sysvar


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_cond = max(4 * threads, 16)
    deck, _tab, tr = config3_setup(n_cond, NODES, seed=20261019)
    tcs = None                                   # the CPU arm has no Tcs evaluation of its own: Tc stands in (no current)
    tr.sweep.jop[:] = 0.0
    if args.warmup > 0:
        oracle_steps(tr, tcs, n_cond, 1, threads, T_START)      # one untimed pass warms caches and the allocator
    _v, total, _x = oracle_steps(tr, tcs, n_cond, args.steps, threads, T_START)
    value = n_cond * NODES * args.steps / total
    sample = (f"{n_cond} conductors x {NODES} nodes per step, {threads} threads "
              "(oracle/props_oracle.c + oracle/step_oracle.c: C port of the reference's Python step path)")
    line = {
        "impl": "reference", "metric": "node_timesteps_per_s", "value": value, "unit": "node-timesteps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(deck.topology),
        "cpu_baseline": {"value": value, "unit": "node-timesteps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "node-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def host_step_sample(device, batch=512, nodes=2001, steps=3):
    """The drop-in data path of `step()` / `step_many()` (opensc2_b200/step.py: pack -> upload -> step -> download ->
    unpack) without the Python attribute packing: every step, ALL step() inputs go host -> device and the solution,
    SYSLOD and the diagnostics come back, through the C ABI with host buffers."""
    from opensc2_b200.batch import BatchedStep
    from opensc2_b200.problem import case1_topology, synthetic_step_arrays

    topo = case1_topology(1)
    arr = synthetic_step_arrays(topo, nodes, batch=batch, seed=5)
    h2d = sum(getattr(arr, f.name).nbytes for f in dataclasses.fields(arr) if isinstance(getattr(arr, f.name), np.ndarray))
    with BatchedStep(topo, batch, nodes, device=device) as bs:
        out = bs.run(arr)                                       # warm
        t0 = time.perf_counter()
        for _ in range(steps):
            out = bs.run(arr)
        sec = time.perf_counter() - t0
    d2h = sum(v.nbytes for v in out.values())
    return {"value": batch * nodes * steps / sec, "unit": "node-timesteps/s", "ms_per_step": 1e3 * sec / steps,
            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "timed_by": "host clock around the calls",
            "workload": f"{batch} CASE_1-shaped conductors x {nodes} nodes, every step() input uploaded and every output "
                        "downloaded each step (BatchedStep.run = osc2_upload_state + osc2_upload_step_inputs + osc2_step + "
                        "osc2_synchronize + osc2_download_state + osc2_download_diagnostics)",
            "pcie_gbs": (h2d + d2h) * steps / sec / 1e9}


def case2_sample(device, batch=1024, steps=3, warmup=3):
    """Secondary workload (BASELINE.json configs[4], CASE_2 half): `batch` CASE_2 variants (NODOFS = 64) on the shipped
    151-node refined mesh through the runner, full device-resident steps, with its own CPU baseline."""
    deck, tab = deck_image("CASE_2_ENEA_HTS_CICC"), synthetic_helium_table()
    rng = np.random.default_rng(7)
    with Transient(deck, [tab], batch, legacy_intial2=True, q0=rng.uniform(100.0, 1000.0, batch), device=device) as tr:
        N, d = tr.zcoord.shape[0], deck.topology.ndf
        tr.time[-1] = 14.5
        tr.run(max_steps=warmup, t_end=1e9)
        tr.bs.set_profiling(True); tr.bs.kernel_ms(reset=True); tr.bs.timer_start()
        for _ in range(steps):
            tr.num_step += 1
            tr.time.append(tr.time[-1] + 0.05)
            tr.bs.advance(tr.time[-1], 0.05, tr.num_step)
        ms = tr.bs.timer_stop(); km = tr.bs.kernel_ms(); tr.bs.synchronize()
        st = tr.bs.download_status()
        threads = os.cpu_count() or 1
        n_cpu = max(threads, 8)
        cpu_val, cpu_sec, _x = oracle_steps(tr, None, n_cpu, 1, threads, 14.5)
    value = batch * N * steps / (ms * 1e-3)
    peak, _src = peaks()
    return {"workload": f"{batch} CASE_2 variants (d={d}: 13 He channels, 25 solids, 97 interfaces) x {N} nodes (refined mesh of the "
                        "deck), dt=0.05 s, full steps (operating conditions + step()), state resident in HBM",
            "value": value, "unit": "node-timesteps/s", "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
            "kernel_ms": {k: v[0] / steps for k, v in km.items()}, "all_status_zero": bool(np.all(st == 0)),
            "alg_bytes_per_node_timestep": b_alg(d), "frac_of_hbm_roofline": value * b_alg(d) / 1e9 / peak,
            "bound": "FP64 pipe and the serial pivot chain (AI 5.3 flop/B, SURVEY.md 8d), not HBM",
            "cpu_baseline": {"value": cpu_val, "unit": "node-timesteps/s", "cores": threads, "kind": "port",
                             "sample": f"{n_cpu} of these conductors x {N} nodes x 1 full step, oracle on {threads} threads ({cpu_sec:.2f} s)"}}


def run_ours(args, rank, world, local_rank):
    import torch

    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    deck, tab, tr = config3_setup(BATCH, NODES, seed=20261019 + rank, device=local_rank)
    topo = deck.topology
    d = topo.ndf
    bs = tr.bs
    bs.synchronize()                              # also reports a conductor whose current is not below Jc(3.5 K)
    tr.time[-1] = T_START
    dt = float(deck.transient["STPMIN"])

    def pinned(shape, dtype):
        t = torch.empty(shape, dtype=dtype, pin_memory=True)
        return t, t.numpy()

    keep_bc, bc_host = pinned(tr.fl_bc.shape, torch.float64)
    bc_host[...] = tr.fl_bc
    keep_diag = [pinned((BATCH, d), torch.float64) for _ in range(3)]
    diag = {"norm_solution": keep_diag[0][1], "norm_change": keep_diag[1][1], "eqteig": keep_diag[2][1]}
    keep_st, status = pinned((BATCH,), torch.int32)

    def barrier():
        bs.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def one_step():
        tr.num_step += 1
        tr.time.append(tr.time[-1] + dt)
        bs.advance(tr.time[-1], dt, tr.num_step)

    # ---- resident: state and parameters already in HBM -------------------------
    for _ in range(args.warmup):
        one_step()
    barrier()
    bs.set_profiling(True)
    bs.kernel_ms(reset=True)
    clocks = ClockSampler(local_rank)
    clocks.start()
    bs.timer_start()
    for _ in range(args.steps):
        one_step()
    ms = bs.timer_stop()
    barrier()
    km = bs.kernel_ms(reset=True)
    bs.set_profiling(False)
    launches = int(sum(v[1] for v in km.values()))

    # ---- end to end through the C ABI: every step uploads that step's boundary values from pinned host memory and
    # reads the step's diagnostics (norms, EQTEIG, status, temperature margin) back -----------------------------
    bs.upload_boundary_values(bc_host); one_step(); bs.download_diagnostics(diag, want_k123=False); bs.download_status(status)   # warm
    barrier()
    margin = None
    bs.timer_start()
    for _ in range(args.steps):
        bs.upload_boundary_values(bc_host)
        one_step()
        bs.download_diagnostics(diag, want_k123=False)
        bs.download_status(status)
        margin = bs.download_margin()
    ms_e2e = bs.timer_stop()
    barrier()
    clk = clocks.stop()
    h2d = bc_host.nbytes
    d2h = sum(v.nbytes for v in diag.values()) + status.nbytes + margin.nbytes
    healthy = bool(np.all(status == 0) and np.all(np.isfinite(diag["norm_solution"])))
    n_total = tr.num_step

    gather = None
    if dist is not None:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        # the only collective of the job: final gather of the results, untimed (here the margins and the norms; the
        # field gather of the full state is exercised by --workload config4)
        from opensc2_b200.shard import any_rank_failed, gather_to_all
        t0 = time.perf_counter()
        full = gather_to_all(np.ascontiguousarray(margin), world * BATCH, device=f"cuda:{local_rank}")
        healthy = healthy and not any_rank_failed(status, device=f"cuda:{local_rank}")
        gather = {"what": "temperature margin of every conductor to every rank (NCCL all_gather), outside the timed region",
                  "bytes": int(full.nbytes), "ms": 1e3 * (time.perf_counter() - t0), "finite": bool(np.all(np.isfinite(full[:, 0])))}
    units = world * BATCH * NODES * args.steps
    value = units / (ms * 1e-3)
    e2e_value = units / (ms_e2e * 1e-3)

    if rank == 0:
        peak, peak_src = peaks()
        alg = kernel_alg_bytes(topo)
        per_step = {k: v[0] / args.steps for k, v in km.items()}          # ms of each kernel family per step
        dom = max((k for k in per_step if k in alg), key=lambda k: per_step[k])
        dom_units = BATCH * (NODES if dom in ("forward", "backward", "norms") else NODES - 1)
        achieved = alg[dom] * dom_units / (per_step[dom] * 1e-3) / 1e9
        step_ms_kernels = sum(per_step[k] for k in per_step if k in alg)
        fracs = {k: alg[k] * BATCH * NODES / (per_step[k] * 1e-3) / 1e9 / peak for k in per_step if k in alg}
        traffic_per_unit, traffic_src = measured_traffic(dom)
        # ---- snapshot of the fields (what a caller pays to look at the solution) and the parity check against the oracle
        t0 = time.perf_counter()
        sv, _sl, _st = bs.download_state(want_syslod=False)
        snap_s = time.perf_counter() - t0
        tcs_g, _tn = bs.download_tcs(nodal=False)
        cpu_threads = os.cpu_count() or 1
        n_cpu = max(2 * cpu_threads, 16)
        cpu_val, cpu_dt, x_cpu = oracle_steps(tr, tcs_g, n_cpu, n_total, cpu_threads, T_START)
        M = topo.n_ch
        a, b = sv[:n_cpu].reshape(n_cpu, NODES, d), x_cpu.reshape(n_cpu, NODES, d)
        err = np.max(np.abs(a - b), axis=1) / np.maximum(np.max(np.abs(b), axis=1), 1e-300)
        parity = {"against": f"CPU oracle loop on conductors 0..{n_cpu - 1} of the timed batch, all {n_total} steps the device took",
                  "max_rel_velocity": float(err[:, :M].max()), "max_rel_pressure_temperature": float(err[:, M:].max()),
                  "tolerance": 2e-6, "ok": bool(err.max() <= 2e-6),
                  "tolerance_note": "the device and the oracle differ in the material fits by 2e-12 (CUDA vs glibc libm); one solve "
                                    "amplifies that by the conditioning of the scaled system (cond ~ 2e8 at 201 nodes, more at "
                                    "10 001): 1e-6 per step in the tests, 2e-6 for this chained check; step() itself on given "
                                    "coefficients is bit-identical (tests/test_gpu_step_parity.py)"}
        line = {
            "metric": "node_timesteps_per_s", "value": value, "unit": "node-timesteps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(topo),
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "node-timesteps/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / args.steps,
                    "call": "BatchedStep.upload_boundary_values + advance + download_diagnostics + download_status + "
                            "download_margin (C ABI, pinned host buffers; state and coefficients stay in HBM)",
                    "solution_finite_and_nonsingular": healthy,
                    "min_temperature_margin_K": float(np.min(margin[:, 0]))},
            "gpu_launches": launches,
            "final_gather": gather,
            "parity_check": parity,
            "snapshot": {"what": "osc2_download_state of the whole batch (solution fields to host)", "bytes": int(sv.nbytes),
                         "ms": 1e3 * snap_s, "gbs": sv.nbytes / snap_s / 1e9},
            "roofline": {
                "bound": "hbm", "kernel": f"osc2_{dom}", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src,
                "alg_bytes_per_unit": alg[dom], "units_per_launch": dom_units, "launch_ms": per_step[dom],
                "traffic": (traffic_per_unit * dom_units) if traffic_per_unit else None,
                "traffic_source": traffic_src,
                "per_kernel_frac": fracs,
                "note": "the forward sweep is a serial dependency chain per conductor (all 4096 conductors are in flight at once: "
                        "time = nodes x cycles per node); DESIGN.md section 5",
                "step": {"alg_bytes_per_node_timestep": b_alg(d), "model": "SURVEY.md 8(d) two-kernel model (assembly + solve)",
                         "achieved": b_alg(d) * BATCH * NODES / (step_ms_kernels * 1e-3) / 1e9,
                         "frac": b_alg(d) * BATCH * NODES / (step_ms_kernels * 1e-3) / 1e9 / peak,
                         "kernel_ms": per_step},
            },
            "cpu_baseline": {"value": cpu_val, "unit": "node-timesteps/s", "cores": cpu_threads, "kind": "port",
                             "sample": f"{n_cpu} conductors x {NODES} nodes x {n_total} full steps of this workload, "
                                       f"oracle/props_oracle.c + oracle/step_oracle.c on {cpu_threads} threads ({cpu_dt:.2f} s)"},
        }
    tr.close()
    if rank == 0:
        if world == 1 and os.environ.get("OSC2_BENCH_EXTRAS", "1") != "0":
            try:
                line["e2e_host_step"] = host_step_sample(local_rank)
            except Exception as exc:      # the headline line must not depend on the secondary samples
                line["e2e_host_step"] = {"error": repr(exc)}
            try:
                line["other_workloads"] = [case2_sample(local_rank)]
            except Exception as exc:
                line["other_workloads"] = [{"workload": "CASE_2", "error": repr(exc)}]
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------
# configs[4]: 131 072 variants, strong scaling
# ---------------------------------------------------------------------------------------------------------
TOTAL4 = int(os.environ.get("OSC2_BENCH_TOTAL4", 131072))
CHUNK2 = int(os.environ.get("OSC2_BENCH_CHUNK2", 8192))        # CASE_2 conductors resident at once (125 GB at 8192)


def run_config4(args, rank, world, local_rank):
    import torch

    from opensc2_b200.shard import shard_range

    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    tab = synthetic_helium_table()
    half = TOTAL4 // 2
    jobs = []                                      # (case, deck, lo, hi) chunks of this rank
    lo1, hi1 = shard_range(half, rank, world)
    jobs.append(("case1", deck_image("CASE_1_ITER_like_LTS"), lo1, hi1))
    lo2, hi2 = shard_range(half, rank, world)
    for c in range(lo2, hi2, CHUNK2):
        jobs.append(("case2", deck_image("CASE_2_ENEA_HTS_CICC"), c, min(c + CHUNK2, hi2)))
    total_ms, node_steps, finals, clk, launches = 0.0, 0, [], None, 0
    per_case_ms = {"case1": 0.0, "case2": 0.0}
    clocks = ClockSampler(local_rank)
    clocks.start()
    for case, deck, lo, hi in jobs:
        n = hi - lo
        rng = np.random.default_rng(1000003 * (1 if case == "case1" else 2) + lo)
        if case == "case1":
            b0 = rng.uniform(0.0, 11.0, n)
            kw = dict(q0=rng.uniform(50.0, 1000.0, n), i_op=rng.uniform(0.0, 60.0e3, n), b_field=np.stack([b0, b0], axis=1))
            t0 = 9.5
        else:
            kw = dict(q0=rng.uniform(100.0, 1000.0, n), legacy_intial2=True)
            t0 = 14.5
        with Transient(deck, [tab], n, device=local_rank, **kw) as tr:
            dt = float(deck.transient["STPMIN"])
            tr.time[-1] = t0

            def one_step():
                tr.num_step += 1
                tr.time.append(tr.time[-1] + dt)
                tr.bs.advance(tr.time[-1], dt, tr.num_step)
            for _ in range(args.warmup):
                one_step()
            tr.bs.synchronize()
            tr.bs.set_profiling(True); tr.bs.kernel_ms(reset=True)
            tr.bs.timer_start()
            for _ in range(args.steps):
                one_step()
            ms = tr.bs.timer_stop()
            launches += int(sum(v[1] for v in tr.bs.kernel_ms().values()))
            tr.bs.synchronize()
            total_ms += ms
            per_case_ms[case] += ms
            node_steps += n * tr.zcoord.shape[0] * args.steps
            sv, _sl, st = tr.bs.download_state(want_syslod=False)
            assert np.all(st == 0), f"{case} chunk {lo}:{hi}: singular conductor"
            finals.append((case, lo, sv))
    clk = clocks.stop()
    # ---- the one collective of the job: final gather of the solution FIELDS of every variant to every rank ---------
    gather = None
    if dist is not None:
        dist.barrier()
        t = torch.tensor([total_ms, float(node_steps)] + [per_case_ms["case1"], per_case_ms["case2"]], dtype=torch.float64, device=f"cuda:{local_rank}")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, node_steps = float(tmax[0]), float(tsum[1])
        per_case_ms = {"case1": float(tmax[2]), "case2": float(tmax[3])}
        t0 = time.perf_counter()
        nbytes = 0
        for case in ("case1", "case2"):
            local = np.concatenate([sv for c, _lo, sv in finals if c == case], axis=0)
            loc = torch.from_numpy(local).to(f"cuda:{local_rank}")
            out = torch.empty((world * loc.shape[0],) + tuple(loc.shape[1:]), dtype=loc.dtype, device=loc.device)
            dist.all_gather_into_tensor(out, loc)          # equal shards: 65 536 is a multiple of 1, 2, 4, 8
            torch.cuda.synchronize()
            nbytes += out.numel() * 8
            ok = bool(torch.isfinite(out).all().item())
            del out, loc
        gather = {"what": "final solution fields (B x Total doubles) of all 131 072 variants to every rank, NCCL all_gather_into_tensor, untimed",
                  "bytes_per_rank": int(nbytes), "ms": 1e3 * (time.perf_counter() - t0), "finite": ok}
    if rank == 0:
        value = node_steps / (total_ms * 1e-3)
        line = {"metric": "node_timesteps_per_s", "value": value, "unit": "node-timesteps/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"BASELINE configs[4]: temperature-margin sweep of {TOTAL4} variants = {half} CASE_1 (d=8, 201 nodes) + "
                                       f"{half} CASE_2 (d=64, 151 nodes, refined mesh), sharded over {world} GPU(s); per rank the CASE_2 share runs "
                                       f"in chunks of <= {CHUNK2} conductors (factor rows: 66 KB per node), chunk after chunk, each chunk's steps "
                                       "timed on the device; handle creation and the initial upload are outside the timed region",
                           "variants": TOTAL4, "chunks_per_rank": len(jobs), "coolant_table": "synthetic helium stand-in"},
                "clocks": clk, "gpu_launches": launches, "final_gather": gather,
                "per_case_ms_per_step": {k: v / args.steps for k, v in per_case_ms.items()},
                "e2e": {"value": value, "unit": "node-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "note": "device-resident sweep: nothing crosses PCIe per step; see the configs[3] line for the host-buffer figures"}}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--workload", default="config3", choices=("config3", "config4"))
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload == "config4":
        run_config4(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
