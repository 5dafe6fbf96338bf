"""The JSON line `bench.py` prints is a contract with the driver: its keys are checked here on the arm that
runs without a GPU (`--impl reference`: the oracle port of the reference's CPU path on the host cores), and the
product arm is checked to fail loudly -- not to fall back to the CPU -- when no CUDA device is visible."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=e,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_line_has_every_contract_key():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line"
    j = json.loads(lines[0])
    assert j["impl"] == "reference"
    assert j["metric"] == "node_timesteps_per_s" and j["unit"] == "node-timesteps/s"
    assert j["n_gpus"] == 1 and j["steps"] == 1 and j["warmup"] == 1
    assert j["higher_is_better"] is True and j["scaling"] == "weak" and j["vs_baseline"] is None
    assert j["dtype"] == "f64" and j["data"] == "synthetic"
    assert "configs[3]" in j["config"]["workload"] and "model" not in j["config"]
    assert j["value"] > 0 and j["ms_per_step"] > 0
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without a CUDA device")
def test_product_arm_refuses_to_run_without_a_gpu():
    r = run_bench("--steps", "1", "--warmup", "1")
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")], "no bench line without the CUDA path"


def test_reference_arm_under_torchrun_prints_one_line_from_rank_zero():
    # the driver launches both arms the same way at N > 1: rank 0 alone runs the CPU reference, the others exit 0
    e = dict(os.environ)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29547", os.path.join(ROOT, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       cwd=ROOT, env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["n_gpus"] == 2 and j["value"] > 0
