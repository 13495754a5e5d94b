"""bench.py's driver contract on the CPU: the reference arm (`--impl reference`, the reference's CPU path timed on
the host cores) prints one JSON line with the agreed keys; ranks other than 0 do no work; the GPU arm has no CPU path."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _run(args, env_extra=None, timeout=280):
    env = dict(os.environ, **(env_extra or {}))
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                          env=env, cwd=str(ROOT))


def test_reference_arm_prints_the_contract_line():
    res = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "LightGlue pairs/sec @2048 kpts" and d["unit"] == "pairs/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and abs(d["value"] - 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]  # one pair per step
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    # "reference": the unmodified module from git-ignored baseline/_ref when that install is present, else the port
    have_ref = (ROOT / "baseline" / "_ref" / "gluefactory").exists()
    assert cb["kind"] == ("reference" if have_ref else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "2048" in cb["sample"]
    # both arms print the same `config` object (the driver compares them)
    import argparse
    sys.path.insert(0, str(ROOT))
    import bench

    ns = argparse.Namespace(workload="c2", pairs=bench.PAIRS_PER_GPU, kpts=bench.KPTS, precision="bf16")
    assert d["config"] == bench.workload_config(ns, 1)
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    res = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
               {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, timeout=120)
    assert res.returncode == 0, res.stderr[-2000:]
    assert not [ln for ln in res.stdout.splitlines() if ln.startswith("{")]


def test_gpu_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("a GPU is present: the arm would run")
    res = _run(["--steps", "1", "--warmup", "1", "--no-cpu-baseline"], timeout=120)
    assert res.returncode != 0
    assert not [ln for ln in res.stdout.splitlines() if ln.startswith("{")]  # no number without the CUDA path
