"""CPU, world_size 2 and 3, gloo: the y-slab halo exchange schedule of cuda_lbm_b200.slab.

A stand-in engine replaces the CUDA handle: its pack_* calls fill the 3*nx-float buffer with a tag encoding
(rank, side, phase, step) and its unpack_* calls record what arrived, so the test can assert that every face
received exactly what the peer's opposite face packed, in the right phase, on odd (neighbour) steps only, and
that CM<OptimalAdapter>'s 3-double all-reduce runs once per step.
"""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cuda_lbm_b200.slab import SlabSolver, neighbours, slab_rows

NX = 16


class FakeEngine:
    def __init__(self, rank, world):
        self.rank, self.world, self.t = rank, world, 0
        self.log = []          # (step, phase, side, received tag)
        self.sums = np.array([1.0 + rank, 2.0, 3.0])
        self.set_sums = []

    def _buf(self, ptr):
        return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)), shape=(3 * NX,))

    def next_step_needs_halo(self):
        return self.world > 1 and ((self.t + 1) & 1) == 1

    def halo(self, what, side, ptr):
        op, phase = what.split("_")
        b = self._buf(ptr)
        if op == "pack":
            b[:] = self.rank * 1000 + side * 100 + (10 if phase == "post" else 0) + (self.t % 10)
        else:
            assert (b == b[0]).all()
            self.log.append((self.t, phase, side, int(b[0])))

    def step(self, n, macroscopics=False):
        assert n == 1
        self.t += 1

    def moment_sums(self):
        return self.sums

    def set_moment_sums(self, s):
        self.set_sums.append(list(s))

    def adapter_prepass(self):
        pass


def _worker(rank, world, port, periodic, optimal, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    e = FakeEngine(rank, world)
    s = SlabSolver(e, NX, periodic, torch.device("cpu"), optimal_adapter=optimal, adapter_exact=True)
    s.step(4)
    q.put((rank, e.log, e.set_sums, s.collectives))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world, periodic, optimal=False):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, periodic, optimal, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = {}
    for _ in range(world):
        r, log, sums, ncoll = q.get(timeout=120)
        res[r] = (log, sums, ncoll)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    return res


@pytest.mark.parametrize("world,periodic", [(2, True), (2, False), (3, True)])
def test_halo_schedule(world, periodic):
    res = _run(world, periodic)
    for rank, (log, _, _) in res.items():
        lo, hi = neighbours(rank, world, periodic)
        peers = (lo, hi)
        nfaces = sum(p is not None for p in peers)
        # 4 steps = 2 odd steps; per odd step one 'pre' and one 'post' message per face
        assert len(log) == 2 * 2 * nfaces
        for (t, phase, side, tag) in log:
            peer = peers[side]
            assert peer is not None
            # the peer packed on ITS side facing me: 1 - side
            # 'pre' is packed before the step (peer's t == mine), 'post' after it (t already advanced on both)
            exp = peer * 1000 + (1 - side) * 100 + (10 if phase == "post" else 0) + (t % 10)
            assert tag == exp, (rank, t, phase, side, tag, exp)
            assert (t % 2 == 0) if phase == "pre" else (t % 2 == 1)


def test_optimal_adapter_allreduce_every_step():
    res = _run(2, True, optimal=True)
    for rank, (_, sums, ncoll) in res.items():
        assert ncoll == 4 and len(sums) == 4
        assert sums[0] == [3.0, 4.0, 6.0]        # (1+0)+(1+1), 2+2, 3+3


def test_slab_rows_cover_the_grid():
    for ny, world in ((32768, 8), (33, 2), (10, 3), (7, 7 // 2)):
        rows = [slab_rows(ny, r, world) for r in range(world)]
        assert rows[0][0] == 0 and sum(n for _, n in rows) == ny
        for (y0, n), (y1, _) in zip(rows, rows[1:]):
            assert y0 + n == y1


def test_neighbours():
    assert neighbours(0, 1, True) == (None, None)
    assert neighbours(0, 4, True) == (3, 1) and neighbours(3, 4, True) == (2, 0)
    assert neighbours(0, 4, False) == (None, 1) and neighbours(3, 4, False) == (2, None)
