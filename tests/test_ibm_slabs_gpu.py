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
