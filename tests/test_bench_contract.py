"""CPU-only checks of bench.py's contract: the reference arm prints ONE JSON line with the agreed keys (and nothing
else on stdout), non-zero ranks of a torchrun launch stay silent, and the CUDA arm refuses to run without a GPU
instead of falling back to a CPU path."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env=None):
    e = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        e.pop(k, None)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, env=e, cwd=ROOT, timeout=600)


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--rollout-steps", "8"])
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                  # one JSON line, nothing else on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "placements/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0
    assert d["metric"].startswith("placements/sec") and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "placements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_reference_arm_other_ranks_are_silent():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cuda_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("checks the no-GPU failure mode")
    r = _run(["--steps", "1", "--warmup", "1"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert not any(ln.startswith("{") for ln in r.stdout.splitlines())     # no number without the CUDA path
