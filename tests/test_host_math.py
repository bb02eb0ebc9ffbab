"""CPU: the collision templates of cuda_lbm_b200/csrc/collide.cuh (the code the CUDA kernels run) compiled as HOST code and checked
against fp64 evaluations of the operators' definitions, for one-cell (V1) and packed two-cell (V2) lanes — tests/host_math_check.cu."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not available")
def test_collision_templates_match_fp64_definitions(tmp_path):
    exe = str(tmp_path / "host_math_check")
    r = subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "-w", "-Xcompiler", "-ffp-contract=off", os.path.join(ROOT, "tests", "host_math_check.cu"), "-o", exe],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:])
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-3000:]
