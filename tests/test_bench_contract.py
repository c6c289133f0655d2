"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) runs the CPU path and prints ONE JSON
line with the keys the driver reads; the product arm refuses to run without CUDA instead of falling back."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT,
                          env=e, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = _run("--impl", "reference", "--steps", "2", "--warmup", "1")
    assert p.returncode == 0, p.stderr
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "heatmaps/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    # "reference" where the reference's own modules are reachable (this container, or the copy shipped under
    # oracle/_ref/reference), else the oracle port
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_runs_the_real_reference_when_it_is_reachable():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("no reference tree or shipped copy")
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    d = json.loads(p.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["kind"] == "reference", d["cpu_baseline"]
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "1", env={"HP_REF_DIR": "/nonexistent"})
    assert p.returncode == 0


@pytest.mark.parametrize("workload", ["regdisp512", "fuse2048"])
def test_reference_arm_of_the_other_workloads(workload):
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--workload", workload)
    assert p.returncode == 0, p.stderr
    d = json.loads(p.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["value"] > 0 and workload[:4] in d["config"]["workload"].lower().replace("-", "") \
        or d["value"] > 0


def test_reference_arm_other_ranks_exit_quietly():
    p = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without CUDA")
def test_product_arm_refuses_to_run_without_cuda():
    p = _run("--steps", "1", "--warmup", "1", "--no-cpu-baseline")
    assert p.returncode != 0
    assert "no CPU fallback" in p.stderr or "CUDA" in p.stderr
