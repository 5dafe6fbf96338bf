#!/usr/bin/env python
"""Benchmark of the batched OPENSC2 time-step path (driver contract in the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric: node-timesteps/s = conductors x nodes x steps / device time, whole job.
Workload (BASELINE.json configs[3]): 4096 CASE_1 variants (swept heat-slug power, background
field, inlet temperature) on a 10 001-node mesh per GPU.  One "step" = one full transient time
step: operating conditions at the Gauss points (coolant tables, correlations, material fits,
heat source; kernel K_P) + step() (assembly, boundary rows, scaling, banded LU, substitution,
norms; kernels K_G, K_F, K_B, K_N).  N > 1: every rank owns its own 4096 conductors (weak
scaling, no collective on the data path).

`--impl reference` times the CPU oracle port of the same work (oracle/props_oracle.c +
oracle/step_oracle.c) on all host threads, rank 0 only: the reference itself is Python and
cannot travel to the GPU box.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from opensc2_b200.model import (case1_model, consistent_initial_state, synthetic_helium_table,  # noqa: E402
                                synthetic_sweep)
from opensc2_b200.problem import case1_topology  # noqa: E402

BATCH = int(os.environ.get("OSC2_BENCH_BATCH", 4096))
NODES = int(os.environ.get("OSC2_BENCH_NODES", 10001))
TILE = 64            # distinct synthetic conductors generated on the host, tiled to the batch
DT = 0.1
T_START = 9.5        # s: the timed steps run through the start of the heat slug (TQBEG = 10 s)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def b_alg(d: int) -> int:
    """algorithmic bytes per node-timestep, two-kernel model."""
    return 8 * (3 * d * (4 * d - 1) + 6 * d)


def kernel_alg_bytes(topo):
    """Per-unit algorithmic bytes of each kernel."""
    d, M, S = topo.ndf, topo.n_ch, topo.n_sol
    nff, nfs, nss, nes = len(topo.ff), len(topo.fs), len(topo.ss), len(topo.es)
    rec = d * (d + 9)
    coef = 1 + 9 * M + 3 * S + 4 * S + 4 * nff + 2 * nfs + 2 * nss + 4 * nes     # per element, K_G inputs
    computed = 9 * M + 3 * S + 4 * S + 2 * nff + nfs + nss + 3 * nes              # per element, K_P outputs
    return {
        "props": 8 * (d + S + computed),                    # per element: state, Tcs in; coefficients out
        "gauss": 8 * (coef + S + rec + 3 * nff),            # per element: coefficients, qsource in; row records, K1..K3 out
        "forward": 8 * (rec + d + 4 * d + d * (2 * d + 2)),  # per node: record, x_old, SYSLOD in/out, factors out
        "backward": 8 * (d * (2 * d + 2) + d),              # per node: factors in, x out
        "norms": 8 * d,                                      # per node: x in
    }


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


def tile(a, batch):
    """Repeat the leading axis of `a` up to `batch`."""
    reps = (batch + a.shape[0] - 1) // a.shape[0]
    return np.ascontiguousarray(np.concatenate([a] * reps, axis=0)[:batch])


def make_workload(batch, nodes, seed):
    """(topology, model, table, StepArrays with geometry/state/boundary values, Sweep), `batch` conductors."""
    import dataclasses

    topo, model, tab = case1_topology(1), case1_model(False), synthetic_helium_table()
    nt = min(TILE, batch)
    small = consistent_initial_state(topo, model, tab, nodes, nt, seed=seed, dt=DT)
    sw = synthetic_sweep(topo, model, nodes, nt, seed=seed)
    keep = ("dz", "fl_bc", "peri_ff", "peri_fs", "peri_ss", "peri_es", "qsource", "sysvar")
    kw = {}
    for f in dataclasses.fields(small):
        v = getattr(small, f.name)
        if isinstance(v, np.ndarray):
            kw[f.name] = tile(v, batch) if f.name in keep else v[:1]      # coefficient arrays are computed on the device
        else:
            kw[f.name] = v
    arr = type(small)(**kw)
    sweep = type(sw)(**{f.name: (tile(getattr(sw, f.name), batch) if getattr(sw, f.name) is not None else None)
                        for f in dataclasses.fields(sw)})
    return topo, model, tab, arr, sweep


def cpu_sample(topo, model, tab, arr, sweep, n_cond, threads, steps=1):
    """Oracle port of one full time step (operating conditions + step()) on `n_cond` conductors."""
    import dataclasses

    from oracle import oracle as O

    O.lib()
    n = arr.n_nodes
    sub = {f.name: getattr(arr, f.name) for f in dataclasses.fields(arr)}
    for k in ("dz", "fl_bc", "peri_ff", "peri_fs", "peri_ss", "peri_es", "qsource", "sysvar"):
        sub[k] = np.ascontiguousarray(sub[k][:n_cond])
    sub["syslod"] = np.zeros((n_cond, n * topo.ndf, 2))
    host = type(arr)(**sub)
    sw = type(sweep)(**{f.name: (getattr(sweep, f.name)[:n_cond] if getattr(sweep, f.name) is not None else None)
                        for f in dataclasses.fields(sweep)})
    t0 = time.perf_counter()
    for k in range(steps):
        ins = O.operating_conditions(topo, model, [tab], sw, host.sysvar, n, T_START + (k + 1) * DT, n_threads=threads)
        for key, v in ins.items():
            setattr(host, key, v)
        host.num_step, host.dt = 2 + k, DT
        out = O.step(topo, host, n_threads=threads)
        host.sysvar, host.syslod = out["sysvar"], out["syslod"]
    dt = time.perf_counter() - t0
    return n_cond * n * steps / dt, dt


def case2_sample(device, batch=1024, nodes=151, steps=3, warmup=2):
    """Secondary workload (BASELINE.json configs[4], CASE_2 half): `batch` CASE_2 variants (NODOFS = 64:
    13 helium channels, 6 HTS + 18 Al strands, SS jacket, 97 interfaces) on the shipped 151-node mesh,
    full device-resident steps through the register-tiled d = 64 sweeps (step_kernels_d64.cuh)."""
    import dataclasses

    from opensc2_b200.batch import BatchedStep
    from opensc2_b200.model import case2_model
    from opensc2_b200.problem import case2_topology

    dt = 0.05
    topo = case2_topology(); model = case2_model(topo); tab = synthetic_helium_table()
    small = consistent_initial_state(topo, model, tab, nodes, 16, seed=1, dt=dt, length=3.0)
    sw = synthetic_sweep(topo, model, nodes, 16, seed=1, length=3.0, with_tcs=False)
    arr = type(small)(**{f.name: (tile(getattr(small, f.name), batch) if isinstance(getattr(small, f.name), np.ndarray)
                                  else getattr(small, f.name)) for f in dataclasses.fields(small)})
    sweep = type(sw)(**{f.name: (tile(getattr(sw, f.name), batch) if getattr(sw, f.name) is not None else None)
                        for f in dataclasses.fields(sw)})
    with BatchedStep(topo, batch, nodes, device=device) as bs:
        bs.set_model(model, [tab]); bs.upload_sweep(sweep); bs.upload_state(arr.sysvar, None); bs.upload_geometry(arr)
        for k in range(warmup):
            bs.advance(11.9 + k * dt, dt, 1 + k)
        bs.synchronize(); bs.set_profiling(True); bs.kernel_ms(reset=True); bs.timer_start()
        for k in range(steps):
            bs.advance(12.0 + k * dt, dt, 1 + warmup + k)
        ms = bs.timer_stop(); km = bs.kernel_ms(); bs.synchronize()
        _sv, _sl, st = bs.download_state()
    d = topo.ndf
    value = batch * nodes * steps / (ms * 1e-3)
    peak, _src = peaks()
    return {"workload": f"{batch} CASE_2 variants (d={d}: 13 He channels, 25 solids, 97 interfaces) x {nodes} nodes, dt={dt} s, "
                        "full steps (operating conditions + step()), state resident in HBM",
            "value": value, "unit": "node-timesteps/s", "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
            "kernel_ms": {k: v[0] / steps for k, v in km.items()}, "all_status_zero": bool(np.all(st == 0)),
            "alg_bytes_per_node_timestep": b_alg(d), "frac_of_hbm_roofline": value * b_alg(d) / 1e9 / peak,
            "bound": "FP64 pipe and the serial pivot chain (AI 5.3 flop/B, SURVEY.md 8d), not HBM"}


def workload_config(topo):
    return {"workload": f"{BATCH} CASE_1 variants per GPU (d={topo.ndf}: 2 He channels + Cu/Nb3Sn strand + SS/glass-epoxy jacket, "
                        f"5 interfaces; swept slug power, field, inlet T), {NODES}-node mesh, Backward Euler dt={DT} s; "
                        "one step = operating conditions at the Gauss points (tables, fits, heat source) + step()",
            "batch_per_gpu": BATCH, "nodes": NODES, "ndf": topo.ndf,
            "l2": "state, coefficients and factors (>100 GB per step) exceed the 126 MB L2, no flush needed"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_cond = max(4 * threads, 16)
    topo, model, tab, arr, sweep = make_workload(n_cond, NODES, seed=20261019)
    if args.warmup > 0:
        cpu_sample(topo, model, tab, arr, sweep, n_cond, threads)      # one untimed pass warms caches and the allocator
    total = 0.0
    for _ in range(args.steps):
        _v, dt = cpu_sample(topo, model, tab, arr, sweep, n_cond, threads)
        total += dt
    value = n_cond * NODES * args.steps / total
    sample = (f"{n_cond} conductors x {NODES} nodes per step, {threads} threads "
              "(oracle/props_oracle.c + oracle/step_oracle.c: C port of the reference's Python step path)")
    line = {
        "impl": "reference", "metric": "node_timesteps_per_s", "value": value, "unit": "node-timesteps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(topo),
        "cpu_baseline": {"value": value, "unit": "node-timesteps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "node-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch

    from opensc2_b200.batch import BatchedStep

    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    topo, model, tab, arr, sweep = make_workload(BATCH, NODES, seed=20261019 + rank)
    d = topo.ndf
    bs = BatchedStep(topo, BATCH, NODES, device=local_rank)
    bs.set_model(model, [tab])
    bs.upload_sweep(sweep)
    bs.upload_state(arr.sysvar, None)           # SYSLOD starts at zero on the device
    bs.upload_geometry(arr)
    bs.synchronize()

    def pinned(shape, dtype):
        t = torch.empty(shape, dtype=dtype, pin_memory=True)
        return t, t.numpy()

    keep_bc, bc_host = pinned(arr.fl_bc.shape, torch.float64)
    bc_host[...] = arr.fl_bc
    keep_diag = [pinned((BATCH, d), torch.float64) for _ in range(3)]
    diag = {"norm_solution": keep_diag[0][1], "norm_change": keep_diag[1][1], "eqteig": keep_diag[2][1]}
    keep_st, status = pinned((BATCH,), torch.int32)

    def barrier():
        bs.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def one_step(k):
        bs.advance(T_START + k * DT, DT, 1 + k)

    # ---- resident: state and parameters already in HBM -------------------------
    n = 0
    for _ in range(args.warmup):
        n += 1
        one_step(n)
    barrier()
    bs.set_profiling(True)
    bs.kernel_ms(reset=True)
    clocks = ClockSampler(local_rank)
    clocks.start()
    bs.timer_start()
    for _ in range(args.steps):
        n += 1
        one_step(n)
    ms = bs.timer_stop()
    barrier()
    km = bs.kernel_ms(reset=True)
    bs.set_profiling(False)
    launches = int(sum(v[1] for v in km.values()))

    # ---- end to end through the C ABI: every step uploads that step's boundary values from
    # pinned host memory and reads the step's diagnostics (norms, EQTEIG, status) back ---------
    n += 1
    bs.upload_boundary_values(bc_host); one_step(n); bs.download_diagnostics(diag); bs.download_status(status)   # warm
    barrier()
    bs.timer_start()
    for _ in range(args.steps):
        n += 1
        bs.upload_boundary_values(bc_host)
        one_step(n)
        bs.download_diagnostics(diag)
        bs.download_status(status)
    ms_e2e = bs.timer_stop()
    barrier()
    clk = clocks.stop()
    h2d = bc_host.nbytes
    d2h = sum(v.nbytes for v in diag.values()) + status.nbytes
    healthy = bool(np.all(status == 0) and np.all(np.isfinite(diag["norm_solution"])))

    gather = None
    if dist is not None:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        # the only collective of the job: final gather of the results (here the solution norms), untimed
        from opensc2_b200.shard import any_rank_failed, gather_to_all
        t0 = time.perf_counter()
        full = gather_to_all(np.ascontiguousarray(diag["norm_solution"]), world * BATCH, device=f"cuda:{local_rank}")
        healthy = healthy and not any_rank_failed(status, device=f"cuda:{local_rank}")
        gather = {"what": "norm_solution of every conductor to every rank (NCCL all_gather), outside the timed region",
                  "bytes": int(full.nbytes), "ms": 1e3 * (time.perf_counter() - t0), "finite": bool(np.all(np.isfinite(full)))}
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
        # measured DRAM traffic per node (read + write) of one launch at THIS size, ncu --set full
        # (profiles/r1_fullsize_ncu_summary.txt)
        traffic_per_unit = {"forward": 3386.0, "gauss": 1464.0, "props": 447.0, "backward": 1216.0}.get(dom)
        cpu_threads = os.cpu_count() or 1
        n_cpu = max(8 * cpu_threads, 16)
        cpu_val, cpu_dt = cpu_sample(topo, model, tab, arr, sweep, n_cpu, cpu_threads)
        line = {
            "metric": "node_timesteps_per_s", "value": value, "unit": "node-timesteps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(topo),
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "node-timesteps/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / args.steps,
                    "call": "BatchedStep.upload_boundary_values + advance + download_diagnostics "
                            "+ download_status (C ABI, pinned host buffers; state and coefficients stay in HBM)",
                    "solution_finite_and_nonsingular": healthy},
            "gpu_launches": launches,
            "final_gather": gather,
            "roofline": {
                "bound": "hbm", "kernel": f"osc2_{dom}", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src,
                "alg_bytes_per_unit": alg[dom], "units_per_launch": dom_units, "launch_ms": per_step[dom],
                "traffic": (traffic_per_unit * dom_units) if traffic_per_unit else None,
                "traffic_source": "ncu --set full dram__bytes_read+write of this kernel at this size (profiles/r1_fullsize_ncu_summary.txt)" if traffic_per_unit else None,
                "per_kernel_frac": fracs,
                "note": "the forward sweep is a serial latency chain per conductor (all 4096 conductors are in flight at once: "
                        "time = nodes x cycles per node), not byte bound; DESIGN.md section 5",
                "step": {"alg_bytes_per_node_timestep": b_alg(d), "model": "SURVEY.md 8(d) two-kernel model (assembly + solve)",
                         "achieved": b_alg(d) * BATCH * NODES / (step_ms_kernels * 1e-3) / 1e9,
                         "frac": b_alg(d) * BATCH * NODES / (step_ms_kernels * 1e-3) / 1e9 / peak,
                         "kernel_ms": per_step},
            },
            "cpu_baseline": {"value": cpu_val, "unit": "node-timesteps/s", "cores": cpu_threads, "kind": "port",
                             "sample": f"{n_cpu} conductors x {NODES} nodes x 1 full step of this workload, "
                                       f"oracle/props_oracle.c + oracle/step_oracle.c on {cpu_threads} threads ({cpu_dt:.2f} s)"},
        }
    bs.close()
    if rank == 0:
        if world == 1 and os.environ.get("OSC2_BENCH_CASE2", "1") != "0":
            try:
                line["other_workloads"] = [case2_sample(local_rank)]
            except Exception as exc:      # the headline line must not depend on the secondary sample
                line["other_workloads"] = [{"workload": "CASE_2", "error": repr(exc)}]
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
