"""bench.py's reference arm (the CPU oracle port, the one leg of the bench that runs without a GPU) prints the contract's
JSON line; under a torchrun-style environment only rank 0 prints and the thread count is not the OMP_NUM_THREADS=1 that
torchrun exports."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "tiny", "--steps", "2",
                           "--warmup", "1", "--gpus", env_extra.get("WORLD_SIZE", "1")], capture_output=True, text=True, env=env, cwd=ROOT,
                          timeout=600)


def test_reference_arm_line():
    r = run({"OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 2 and line["warmup"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    assert line["cpu_baseline"]["cores"] == cores
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "sample" in line["config"] and "workload" in line["config"]


def test_reference_arm_other_ranks_are_silent():
    r = run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""
