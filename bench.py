#!/usr/bin/env python
"""Benchmark of the batched OPENSC2 time-step path (driver contract in the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric: node-timesteps/s = conductors x nodes x steps / device time, whole job.
Workload (BASELINE.json configs[3]): 4096 CASE_1 variants on a 10 001-node mesh
per GPU (weak scaling for N > 1: every rank owns its own 4096 conductors, no
data-path collective).  `--impl reference` times the CPU oracle port of the
reference's step() (oracle/step_oracle.c) on all host threads, rank 0 only.
"""
from __future__ import annotations

import argparse
import dataclasses
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from opensc2_b200.problem import case1_topology, synthetic_step_arrays  # noqa: E402

BATCH = int(os.environ.get("OSC2_BENCH_BATCH", 4096))
NODES = int(os.environ.get("OSC2_BENCH_NODES", 10001))
TILE = 64            # distinct synthetic conductors generated on the host, tiled to the batch
DT = 0.1


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def b_alg(d: int) -> int:
    """algorithmic bytes per node-timestep, two-kernel model."""
    return 8 * (3 * d * (4 * d - 1) + 6 * d)


def kernel_alg_bytes(d: int, M: int, S: int, n_coef: int):
    """Per-unit algorithmic bytes of each kernel of the fast path."""
    rec = d * (d + 9)
    return {
        "gauss": 8 * (n_coef + rec),                       # per element: coefficients in, row records out
        "forward": 8 * (rec + d + 4 * d + d * (2 * d + 2)),  # per node: record, x_old, SYSLOD in/out, factors out
        "backward": 8 * (d * (2 * d + 2) + d),             # per node: factors in, x out
        "norms": 8 * d,                                     # per node: x in
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
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
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


def tiled_inputs(topo, batch, nodes, seed, pinned=False):
    """`TILE` distinct synthetic conductors repeated up to `batch`; optionally written
    straight into page-locked host buffers (torch is used as the allocator only)."""
    small = synthetic_step_arrays(topo, nodes, batch=min(TILE, batch), seed=seed, dt=DT)
    nt = small.dz.shape[0]
    keep, kw = [], {}
    for f in dataclasses.fields(small):
        v = getattr(small, f.name)
        if not isinstance(v, np.ndarray):
            kw[f.name] = v
            continue
        shape = (batch,) + v.shape[1:]
        if pinned:
            import torch
            t = torch.empty(shape, dtype=torch.float64, pin_memory=True)
            keep.append(t)
            out = t.numpy()
        else:
            out = np.empty(shape)
        for lo in range(0, batch, nt):
            hi = min(batch, lo + nt)
            out[lo:hi] = v[:hi - lo]
        kw[f.name] = out
    arr = type(small)(**kw)
    arr.__dict__["_pins"] = keep
    return arr


def cpu_sample(topo, arr, n_cond, threads, steps=1):
    """Oracle port of the reference step() on `n_cond` conductors of the workload."""
    from oracle import oracle as O

    kw = {}
    for f in dataclasses.fields(arr):
        v = getattr(arr, f.name)
        kw[f.name] = np.ascontiguousarray(v[:n_cond]) if isinstance(v, np.ndarray) else v
    sub = type(arr)(**kw)
    O.lib()
    t0 = time.perf_counter()
    for k in range(steps):
        sub.num_step = 2 + k
        out = O.step(topo, sub, n_threads=threads)
        sub.sysvar, sub.syslod = out["sysvar"], out["syslod"]
    dt = time.perf_counter() - t0
    return n_cond * arr.n_nodes * steps / dt, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    topo = case1_topology(1)
    threads = os.cpu_count() or 1
    n_cond = max(threads, 16)
    arr = tiled_inputs(topo, n_cond, NODES, seed=20261019)
    if args.warmup > 0:
        cpu_sample(topo, arr, n_cond, threads)      # one untimed pass warms caches and the allocator
    times = []
    for _ in range(args.steps):
        v, dt = cpu_sample(topo, arr, n_cond, threads)
        times.append(dt)
    total = sum(times)
    value = n_cond * NODES * args.steps / total
    sample = f"{n_cond} conductors x {NODES} nodes per step, {threads} threads (oracle/step_oracle.c)"
    line = {
        "impl": "reference", "metric": "node_timesteps_per_s", "value": value, "unit": "node-timesteps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(topo),
        "cpu_baseline": {"value": value, "unit": "node-timesteps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "node-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(topo):
    return {"workload": f"{BATCH} CASE_1 variants per GPU (d={topo.ndf}: 2 He channels + strand + jacket, 5 interfaces), "
                        f"{NODES}-node mesh, Backward Euler dt={DT} s, step() on Gauss-point coefficients",
            "batch_per_gpu": BATCH, "nodes": NODES, "ndf": topo.ndf,
            "l2": "inputs and factors (>100 GB per step) exceed the 126 MB L2, no flush needed"}


def run_ours(args, rank, world, local_rank):
    import torch

    from opensc2_b200.batch import BatchedStep

    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    topo = case1_topology(1)
    d = topo.ndf
    arr = tiled_inputs(topo, BATCH, NODES, seed=20261019 + rank, pinned=True)
    h2d = sum(getattr(arr, n).nbytes for n in ("dz", "fl_gauss", "fl_node", "fl_bc", "sol_gauss", "sol_q", "peri_ff",
                                                "htc_ff", "peri_fs", "htc_fs", "peri_ss", "htc_ss", "peri_es", "htc_es", "qsource"))
    n_coef = h2d // (8 * BATCH * (NODES - 1))
    bs = BatchedStep(topo, BATCH, NODES, device=local_rank)
    bs.upload_state(arr.sysvar, arr.syslod)
    bs.upload_step_inputs(arr)
    bs.synchronize()

    def barrier():
        bs.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- resident: inputs already in HBM -----------------------------------
    for k in range(args.warmup):
        bs.step(DT, 2 + k)
    barrier()
    bs.set_profiling(True)
    bs.kernel_ms(reset=True)
    clocks = ClockSampler(local_rank)
    clocks.start()
    bs.timer_start()
    for k in range(args.steps):
        bs.step(DT, 2 + args.warmup + k)
    ms = bs.timer_stop()
    barrier()
    clk = clocks.stop()
    km = bs.kernel_ms(reset=True)
    bs.set_profiling(False)
    launches = int(sum(v[1] for v in km.values()))

    # ---- end to end: host buffers in, diagnostics out, every step ----------
    diag_bytes = 3 * BATCH * d * 8 + 4 * BATCH
    bs.upload_step_inputs(arr); bs.step(DT, 2); bs.download_diagnostics()      # warm
    barrier()
    bs.timer_start()
    for k in range(args.steps):
        bs.upload_step_inputs(arr)
        bs.step(DT, 3 + k)
        bs.download_diagnostics()
        bs.download_state(want_syslod=False)[2]
    ms_e2e = bs.timer_stop()
    barrier()
    d2h = diag_bytes + BATCH * d * NODES * 8

    if dist is not None:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    units = world * BATCH * NODES * args.steps
    value = units / (ms * 1e-3)
    e2e_value = units / (ms_e2e * 1e-3)

    if rank == 0:
        peak, peak_src = peaks()
        alg = kernel_alg_bytes(d, topo.n_ch, topo.n_sol, n_coef)
        per_launch = {k: (v[0] / max(v[1], 1)) for k, v in km.items()}
        dom = max((k for k in per_launch if k in alg), key=lambda k: per_launch[k])
        dom_units = BATCH * (NODES - 1 if dom == "gauss" else NODES)
        achieved = alg[dom] * dom_units / (per_launch[dom] * 1e-3) / 1e9
        step_ms_kernels = sum(per_launch[k] for k in per_launch if k in alg)
        # measured DRAM traffic per node of the forward sweep (profiles/): read + write, ncu --set full
        traffic_per_unit = {"forward": 2162.0}.get(dom)
        cpu_val, cpu_dt = cpu_sample(topo, arr, max(os.cpu_count() or 1, 16), os.cpu_count() or 1)
        line = {
            "metric": "node_timesteps_per_s", "value": value, "unit": "node-timesteps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(topo),
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "node-timesteps/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / args.steps,
                    "call": "BatchedStep.upload_step_inputs + step + download_diagnostics + download_state (C ABI, pinned host buffers)"},
            "gpu_launches": launches,
            "roofline": {
                "bound": "hbm", "kernel": f"osc2_{dom}", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src,
                "alg_bytes_per_unit": alg[dom], "units_per_launch": dom_units, "launch_ms": per_launch[dom],
                "traffic": (traffic_per_unit * dom_units) if traffic_per_unit else None,
                "traffic_source": "ncu dram__bytes_read+write per node at N=201 (profiles/), scaled to this launch" if traffic_per_unit else None,
                "step": {"alg_bytes_per_node_timestep": b_alg(d), "model": "SURVEY.md 8(d) two-kernel model",
                         "achieved": b_alg(d) * BATCH * NODES / (step_ms_kernels * 1e-3) / 1e9,
                         "frac": b_alg(d) * BATCH * NODES / (step_ms_kernels * 1e-3) / 1e9 / peak,
                         "kernel_ms": per_launch},
            },
            "cpu_baseline": {"value": cpu_val, "unit": "node-timesteps/s", "cores": os.cpu_count() or 1, "kind": "port",
                             "sample": f"{max(os.cpu_count() or 1, 16)} conductors x {NODES} nodes x 1 step of this workload, "
                                       f"oracle/step_oracle.c on {os.cpu_count()} threads ({cpu_dt:.2f} s)"},
        }
        print(json.dumps(line), flush=True)
    bs.close()
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
