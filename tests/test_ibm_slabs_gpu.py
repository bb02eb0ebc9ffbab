"""GPU: immersed bodies on a y-slab decomposed domain — bodies whose marker stencils cross a slab
face, several bodies, overlapping bodies.  Bar: the slab-decomposed run equals the single-handle run BIT FOR BIT (every slab
of a body runs the same iterations on the same node states), with the halo coupling (lbm_ibm_pack -> sum over slabs ->
lbm_ibm_unpack, what slab.py does with an all-reduce) and with peer-mapped neighbours (node states stored into the
neighbours' mailboxes by the pre-pass kernel); and the single-handle run equals the CPU oracle to fp32 round-off.
"""
import numpy as np
import pytest

import cases
from cases import make_engine, make_oracle
from test_parity_gpu import TOL_F, TOL_RHO, _run_slabs

pytestmark = pytest.mark.gpu


def _cyl(cx, cy, r, n):
    from oracle import oracle as O
    return O.create_cylinder(cx, cy, r, n)


def _case(name, coll, bodies, nx=96, ny=48):
    c = cases.Case(name, nx, ny, coll, cases._cyl_nu(ny), (False, False), 0.05, "cyl_ibm")
    c.bodies = bodies
    return c


def _single(case, nsteps):
    e = make_engine(case)
    e.init_fields(*case.init_fields())
    e.step(nsteps, macroscopics=True)
    out = e.macroscopics(), e.populations(), e.info()
    e.close()
    return out


# the reference cylinder (centre row 24, radius 3: stencil rows 21..28) on 96x48
ONE = lambda: [_cyl(18.0, 24.0, 3.0, 16)]
# + a second one far away in the top rows + a third one overlapping the first (shares lattice nodes -> one group)
THREE = lambda: [_cyl(18.0, 24.0, 3.0, 16), _cyl(60.3, 40.2, 2.5, 12), _cyl(21.5, 25.5, 2.0, 10)]
# a large body across three slabs of 8 rows (rows 15..34)
BIG = lambda: [_cyl(30.0, 24.5, 9.0, 48)]


@pytest.mark.parametrize("bodies,world,coll", [(ONE, 2, cases.MRT), (ONE, 3, cases.BGK), (ONE, 4, cases.CM), (THREE, 2, cases.MRT),
                                               (THREE, 4, cases.BGK), (BIG, 6, cases.MRT), (THREE, 2, cases.CM_OPT)])
def test_halo_coupled_slabs_with_bodies_match_single_domain(bodies, world, coll):
    case = _case("ibm_slabs", coll, bodies())
    nsteps = 7
    rho_s, u_s, f_s = _run_slabs(case, world, nsteps)
    (rho_1, u_1), f_1, info = _single(case, nsteps)
    assert info.num_markers == sum(len(b) for b in case.bodies) and info.num_ibm_nodes > 0
    tol = 0.0 if coll != cases.CM_OPT else 2e-7
    assert np.isfinite(f_1).all()
    assert np.abs(f_s - f_1).max() <= tol, np.abs(f_s - f_1).max()
    assert np.abs(rho_s - rho_1).max() <= tol and np.abs(u_s - u_1).max() <= tol


@pytest.mark.parametrize("bodies,world,coll,chunk", [(ONE, 2, cases.MRT, 7), (ONE, 4, cases.BGK, 3), (THREE, 2, cases.CM, 2), (THREE, 4, cases.MRT, 7),
                                                     (THREE, 3, cases.CM_OPT, 1)])
def test_peer_mapped_slabs_with_bodies_match_single_domain(bodies, world, coll, chunk):
    case = _case("ibm_slabs_direct", coll, bodies())
    nsteps = 7
    rho_s, u_s, f_s = _run_slabs(case, world, nsteps, direct=True, chunk=chunk)
    (rho_1, u_1), f_1, _ = _single(case, nsteps)
    tol = 0.0 if coll != cases.CM_OPT else 2e-7
    assert np.abs(f_s - f_1).max() <= tol, np.abs(f_s - f_1).max()
    assert np.abs(rho_s - rho_1).max() <= tol and np.abs(u_s - u_1).max() <= tol


@pytest.mark.parametrize("bodies,world,coll,chunk", [(BIG, 6, cases.MRT, 7), (THREE, 5, cases.CM_OPT, 3), (BIG, 4, cases.BGK, 2)])
def test_all_mapped_slabs_carry_a_body_over_any_number_of_slabs(bodies, world, coll, chunk):
    """lbm_peer_attach_all: the owner of a stencil node stores its state into EVERY slab's mailbox, a slab that works on a body
    waits for the stage counters of all the slabs that own part of it — BIG spans four of the six 8-row slabs."""
    case = _case("ibm_slabs_all", coll, bodies())
    nsteps = 7
    rho_s, u_s, f_s = _run_slabs(case, world, nsteps, direct="all", chunk=chunk)
    (rho_1, u_1), f_1, _ = _single(case, nsteps)
    tol = 0.0 if coll != cases.CM_OPT else 2e-7
    assert np.abs(f_s - f_1).max() <= tol, np.abs(f_s - f_1).max()
    assert np.abs(rho_s - rho_1).max() <= tol and np.abs(u_s - u_1).max() <= tol


def test_body_over_three_slabs_is_refused_by_the_peer_mapped_coupling():
    """Peer-mapped slabs reach their two neighbours only: the outer slabs of a three-slab body cannot see the far one."""
    import cuda_lbm_b200 as L
    case = _case("ibm_big", cases.BGK, BIG())
    world = 6
    engs = [make_engine(case, rank=r, world=world) for r in range(world)]
    for e in engs:
        e.init_fields(*case.init_fields())
    descs = [e.peer_export() for e in engs]
    for r, e in enumerate(engs):
        if r > 0:
            e.peer_attach(0, descs[r - 1])
        if r < world - 1:
            e.peer_attach(1, descs[r + 1])
    with pytest.raises(L.LbmError) as ei:       # slab 1 owns row 15 of the body, whose last rows belong to slab 4
        engs[1].step(1)
    assert "halo coupling" in str(ei.value)
    for e in engs:
        e.close()


def test_mailbox_capacity_is_checked():
    import cuda_lbm_b200 as L
    case = _case("ibm_cap", cases.BGK, [])
    e = make_engine(case, rank=0, world=2, ibm_mailbox_nodes=8)
    with pytest.raises(L.LbmError) as ei:
        e.add_body(_cyl(18.0, 24.0, 3.0, 16))
    assert "mailbox" in str(ei.value)
    assert e.info().num_markers == 0            # the failed call left no half-added body behind
    e.add_body(_cyl(18.0, 24.0, 0.4, 2))        # 2 markers in one cell: 4 nodes fit
    assert e.info().num_markers == 2
    e.close()


def test_halo_coupling_demands_the_exchange():
    import cuda_lbm_b200 as L
    case = _case("ibm_need", cases.BGK, ONE())
    engs = [make_engine(case, rank=r, world=2) for r in range(2)]
    for e in engs:
        e.init_fields(*case.init_fields())
    assert engs[0].ibm_exchange_floats() == engs[1].ibm_exchange_floats() > 0
    with pytest.raises(L.LbmError) as ei:
        engs[0].step(1)
    assert "lbm_ibm_pack" in str(ei.value)
    for e in engs:
        e.close()


@pytest.mark.parametrize("coll", [cases.BGK, cases.MRT])
def test_several_bodies_match_oracle(coll):
    """Single handle, three bodies added one by one (two of them overlapping) against the oracle's one marker array."""
    case = _case("ibm_three", coll, THREE())
    rho0, u0 = case.init_fields()
    o, e = make_oracle(case), make_engine(case)
    o.init(rho0, u0); e.init_fields(rho0, u0)
    done = 0
    for n in (1, 2, 3, 10, 40):
        o.step(n - done); e.step(n - done, macroscopics=True); done = n
        df = np.abs(e.populations() - o.populations()).max()
        dr = np.abs(e.macroscopics()[0] - o.macroscopics()[0]).max()
        assert df <= TOL_F * n ** 0.5 and dr <= TOL_RHO * n ** 0.5, (coll, n, df, dr)
    e.close()


# ---------------------------------------------------------------- IBMBody::velocities and moving bodies (SURVEY A-D9, §8f-2)
def _spin(points, centre, omega):
    """rigid rotation about `centre`: the velocity of every marker"""
    d = points - np.asarray(centre, np.float32)
    return np.stack([-omega * d[:, 1], omega * d[:, 0]], axis=1).astype(np.float32)


@pytest.mark.parametrize("quirks", [63, 0])
def test_marker_velocities_match_oracle(quirks):
    """With LBM_QK_D9_IBM_ZERO_TARGET clear the markers force the fluid towards IBMBody::velocities (a spinning cylinder);
    engine and oracle agree to fp32 round-off, with the reference's other IBM defects on (63) and repaired (0)."""
    body = _cyl(30.0, 24.0, 6.0, 40)
    vel = _spin(body, (30.0, 24.0), 0.004)
    case = _case("ibm_spin", cases.MRT, [body])
    rho0, u0 = case.init_fields()
    o, e = make_oracle(case, quirks), make_engine(case, quirks)
    o.set_marker_velocities(vel); e.set_body_velocities(0, vel)
    o.init(rho0, u0); e.init_fields(rho0, u0)
    done = 0
    for n in (1, 2, 3, 10, 40):
        o.step(n - done); e.step(n - done, macroscopics=True); done = n
        df = np.abs(e.populations() - o.populations()).max()
        dr = np.abs(e.macroscopics()[0] - o.macroscopics()[0]).max()
        assert df <= TOL_F * n ** 0.5 and dr <= TOL_RHO * n ** 0.5, (quirks, n, df, dr)
    # the target velocities do something: the same run without them differs
    e2 = make_engine(case, quirks)
    e2.init_fields(rho0, u0)
    e2.step(40)
    assert np.abs(e2.populations() - e.populations()).max() > 1e-6
    e.close(); e2.close()


def test_marker_velocities_are_dead_data_with_the_reference_quirk():
    """LBM_QK_REFERENCE (bit 64 set) reproduces the reference: velocities are carried and ignored (IBM_impl.cuh:15)."""
    body = _cyl(30.0, 24.0, 6.0, 40)
    case = _case("ibm_dead", cases.BGK, [body])
    rho0, u0 = case.init_fields()
    outs = []
    for with_vel in (False, True):
        e = make_engine(case, 127)
        if with_vel:
            e.set_body_velocities(0, _spin(body, (30.0, 24.0), 0.004))
        e.init_fields(rho0, u0)
        e.step(12)
        outs.append(e.populations())
        e.close()
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("world", [1, 2])
def test_moved_body_matches_oracle_and_slabs(world):
    """lbm_move_body between steps (the body crosses the slab face while moving) against the oracle with re-set markers;
    the two-slab run equals the single-handle run bit for bit."""
    import torch
    b0 = _cyl(30.0, 21.0, 4.0, 24)
    other = _cyl(70.0, 30.0, 3.0, 12)
    case = _case("ibm_move", cases.MRT, [b0, other])
    rho0, u0 = case.init_fields()
    shifts = [(0.0, 0.0), (0.37, 1.21), (0.74, 2.42), (1.11, 3.63)]
    o = make_oracle(case)
    o.init(rho0, u0)
    engs = [make_engine(case, rank=r, world=world) for r in range(world)]
    for e in engs:
        e.init_fields(rho0, u0)
    halo = {(r, s): torch.zeros(3 * case.nx, dtype=torch.float32, device="cuda") for r in range(world) for s in (0, 1)}

    def step_all():
        need = engs[0].next_step_needs_halo()
        for phase in (("pre",) if need else ()):
            _halo(phase)
        nf = engs[0].ibm_exchange_floats()          # the moved body touches a different number of lattice nodes
        assert all(e.ibm_exchange_floats() == nf for e in engs)
        if nf:
            bufs = [torch.zeros(nf, dtype=torch.float32, device="cuda") for _ in engs]
            torch.cuda.synchronize()                # torch works on its own stream, the handles on theirs
            for e, b in zip(engs, bufs):
                e.ibm_pack(b.data_ptr())
            for e in engs:
                e.sync()
            tot = torch.stack(bufs).sum(dim=0)
            torch.cuda.synchronize()
            for e in engs:
                e.ibm_unpack(tot.data_ptr())
        for e in engs:
            e.step(1, macroscopics=True)
        for e in engs:
            e.sync()
        if need:
            _halo("post")

    def _halo(phase):
        # two slabs, non-periodic: slab 0's upper face (side 1) meets slab 1's lower face (side 0)
        engs[0].halo("pack_" + phase, 1, halo[(0, 1)].data_ptr()); engs[1].halo("pack_" + phase, 0, halo[(1, 0)].data_ptr())
        for e in engs:
            e.sync()
        engs[0].halo("unpack_" + phase, 1, halo[(1, 0)].data_ptr()); engs[1].halo("unpack_" + phase, 0, halo[(0, 1)].data_ptr())
        for e in engs:
            e.sync()

    for (dx, dy) in shifts:
        moved = (b0 + np.array([dx, dy], np.float32)).astype(np.float32)
        o.set_markers(np.concatenate([moved, other], axis=0))
        for e in engs:
            e.move_body(0, moved)
        for _ in range(4):
            o.step(1)
            step_all()
    f = np.concatenate([e.populations() for e in engs], axis=0)
    rho = np.concatenate([e.macroscopics()[0] for e in engs], axis=0)
    for e in engs:
        e.close()
    n = 4 * len(shifts)
    assert np.abs(f - o.populations()).max() <= TOL_F * n ** 0.5 and np.abs(rho - o.macroscopics()[0]).max() <= TOL_RHO * n ** 0.5
    if world == 1:
        test_moved_body_matches_oracle_and_slabs.single = f
    elif getattr(test_moved_body_matches_oracle_and_slabs, "single", None) is not None:
        assert np.array_equal(f, test_moved_body_matches_oracle_and_slabs.single)


def test_body_index_is_checked():
    import cuda_lbm_b200 as L
    case = _case("ibm_idx", cases.BGK, ONE())
    e = make_engine(case)
    for call in (lambda: e.set_body_velocities(1, np.zeros((16, 2), np.float32)), lambda: e.move_body(-1, np.zeros((16, 2), np.float32))):
        with pytest.raises(L.LbmError):
            call()
    e.close()


@pytest.mark.parametrize("coll,nsteps", [(cases.MRT, 9), (cases.BGK, 40)])
def test_many_block_ibm_path_is_bit_identical(coll, nsteps, monkeypatch):
    """Bodies with more markers / stencil nodes than one block handles quickly (LBM_B200_IBM_ONE_BLOCK_MAX, default 8192) take seven
    many-block launches per step instead of one block with barriers: same per-marker and per-node arithmetic, same bits —
    also inside the replayed CUDA graphs (40 steps = two 16-step graphs + 8 launches)."""
    case = _case("ibm_blocks", coll, THREE())
    rho0, u0 = case.init_fields()
    outs, launches = [], []
    for limit in ("8192", "0"):
        monkeypatch.setenv("LBM_B200_IBM_ONE_BLOCK_MAX", limit)
        e = make_engine(case)
        e.init_fields(rho0, u0)
        l0 = e.info().kernel_launches
        e.step(nsteps, macroscopics=True)
        outs.append((e.populations(), e.macroscopics()))
        launches.append(e.info().kernel_launches - l0)
        e.close()
    assert launches[1] == launches[0] + 6 * nsteps            # 7 launches instead of 1, every step
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1][1], outs[1][1][1])
