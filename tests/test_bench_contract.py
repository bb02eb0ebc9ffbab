"""CPU: the reference arm of bench.py (`--impl reference`: the CPU restatement of the reference's path on the host cores) prints one JSON
line with the contract's keys, on rank 0 only, and sets its OpenMP thread count itself (torchrun exports OMP_NUM_THREADS=1)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                          capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_reference_arm_prints_the_contract_line_with_all_host_threads():
    r = _run({"OMP_NUM_THREADS": "1"})          # what torchrun hands to its workers
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "MLUPS" and d["unit"] == "MLUPS" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("taylor_green_d2q9_bgk_32768x32768")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] == min(os.cpu_count(), cb["cores"]) and cb["cores"] >= 1
    if os.cpu_count() > 1:
        assert cb["cores"] > 1, "the arm must not inherit OMP_NUM_THREADS=1"
    assert d["e2e"] == {"value": d["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
