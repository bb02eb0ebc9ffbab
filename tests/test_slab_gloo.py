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
    def __init__(self, rank, world, ibm_floats=0):
        self.rank, self.world, self.t = rank, world, 0
        self.ibm_floats, self.ibm_log = ibm_floats, []
        self.calls = []
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
        self.calls.append(("prepass", self.t))

    def adapter_sums_pending(self):
        """lagged mode: only the first step after init has no sums from a previous step"""
        return self.t == 0

    # driver segment from / to host memory (SlabSolver.run_from_host, halo coupling = the three calls with a barrier)
    def init_fields_local(self, rho_ptr, u_ptr):
        self.calls.append(("init", rho_ptr, u_ptr))
        self.t = 0

    def sync(self):
        self.calls.append(("sync",))

    def macroscopics_into(self, rho_ptr, u_ptr):
        self.calls.append(("read", self.t, rho_ptr, u_ptr))

    # bodies across slab faces: slab r "owns" every world-th node state, the rest of its buffer is zero
    def ibm_exchange_floats(self):
        return self.ibm_floats

    def _ibm_buf(self, ptr):
        return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)), shape=(self.ibm_floats,))

    def ibm_pack(self, ptr):
        b = self._ibm_buf(ptr)
        b[:] = 0
        b[self.rank::self.world] = 100 * (self.rank + 1) + self.t

    def ibm_unpack(self, ptr):
        self.ibm_log.append((self.t, self._ibm_buf(ptr).copy()))


def _worker(rank, world, port, periodic, optimal, q, ibm_floats=0):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    e = FakeEngine(rank, world, ibm_floats if ibm_floats != -2 else 0)
    s = SlabSolver(e, NX, periodic, torch.device("cpu"), optimal_adapter=optimal, adapter_exact=(ibm_floats != -2))
    if ibm_floats == -2:            # OptimalAdapter with the grid means of the previous step (LBM_ADAPTER_LAGGED)
        s.step(4)
        q.put((rank, e.calls, e.set_sums, s.collectives))
        dist.barrier()
        dist.destroy_process_group()
        return
    if ibm_floats == -1:            # the driver-segment call instead of plain stepping
        e.ibm_floats = 0
        s.run_from_host(11, 12, 4, 13, 14)
        q.put((rank, e.log, e.calls, s.collectives))
        dist.barrier()
        dist.destroy_process_group()
        return
    s.step(4)
    if ibm_floats:
        q.put((rank, e.ibm_log, None, s.collectives))
    else:
        q.put((rank, e.log, e.set_sums, s.collectives))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world, periodic, optimal=False, ibm_floats=0):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, periodic, optimal, q, ibm_floats)) for r in range(world)]
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


def test_lagged_optimal_adapter_bootstraps_the_first_step():
    """LBM_ADAPTER_LAGGED on several slabs with the halo coupling: the sums come out of each step (one all-reduce after it), except before
    the very first step, which has no predecessor: a moments pre-pass + all-reduce bootstraps it (5 collectives for 4 steps)."""
    res = _run(2, True, optimal=True, ibm_floats=-2)
    for rank, (calls, sums, ncoll) in res.items():
        assert [c for c in calls if c[0] == "prepass"] == [("prepass", 0)], calls
        assert ncoll == 5 and len(sums) == 5


@pytest.mark.parametrize("world", [2, 3])
def test_ibm_node_states_are_gathered_by_allreduce_every_step(world):
    """Bodies across slab faces with the halo coupling: before EVERY step each slab packs the node states it owns (zeros
    elsewhere), the sum over slabs hands every slab the complete set, bit for bit (x + 0 = x)."""
    nf = 10
    res = _run(world, False, ibm_floats=nf)
    for rank, (log, _, ncoll) in res.items():
        assert ncoll == 4 and [t for t, _ in log] == [0, 1, 2, 3]
        for t, got in log:
            want = np.zeros(nf, np.float32)
            for r in range(world):
                want[r::world] = 100 * (r + 1) + t
            assert np.array_equal(got, want), (rank, t, got, want)


def test_run_from_host_with_the_halo_coupling_is_init_barrier_steps_readback():
    """SlabSolver.run_from_host without peer-mapped neighbours: lbm_init_fields_local, sync + barrier, the stepped halo schedule
    (2 odd steps of 4 -> a 'pre' and a 'post' message per face each), then the read-back of step 4 into the caller's buffers."""
    res = _run(2, True, ibm_floats=-1)
    for rank, (log, calls, _) in res.items():
        assert calls[0] == ("init", 11, 12) and calls[1] == ("sync",) and calls[-1] == ("read", 4, 13, 14), calls
        assert len(log) == 2 * 2 * 2


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
