"""GPU, 2+ devices: the y-slab path across PROCESSES (one rank per GPU, torchrun) — CUDA IPC peer mapping over NVLink, NCCL
halo rows, IBM node states — against the single-GPU run, bit for bit.  Skipped on a one-GPU box (the same schedule is
covered there by several handles on one GPU, tests/test_parity_gpu.py and tests/test_ibm_slabs_gpu.py, and on CPU by
tests/test_slab_gloo.py)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("kind,mode,coll,extra", [("tg", "direct", 0, ()), ("tg", "nccl", 1, ()), ("ibm", "direct", 1, ()), ("ibm", "nccl", 1, ()),
                                                  ("ibm", "direct", 3, ()), ("ibm", "nccl", 3, ()), ("tg", "direct", 1, ("--from-host",)),
                                                  ("tg", "nccl", 0, ("--from-host",)),
                                                  ("tg", "nccl", 3, ("--adapter", "1")), ("tg", "direct", 3, ("--adapter", "1")), ("ibm", "direct", 3, ("--adapter", "1"))])
def test_slabs_across_processes_match_single_gpu(kind, mode, coll, extra):
    n = min(_ngpu(), 4)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_port()), os.path.join(HERE, "mp_slab_worker.py"), "--kind", kind, "--mode", mode, "--coll", str(coll), *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "MP_PARITY" in r.stdout and " OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
