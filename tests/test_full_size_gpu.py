"""GPU: the five BASELINE.json configurations at their FULL sizes.

Configs 1 and 2 are small enough for the CPU oracle, so they are compared with it directly.  Configs 3-5 are checked through
properties that do not depend on the size: invariance under the y-slab decomposition (bit for bit, the slabs being handles on
the same GPU, peer-mapped), the analytic Taylor-Green field (error sums taken on the device), mass conservation
with the stale-rest-population defect repaired, finiteness.
"""
import numpy as np
import pytest

import cases
from cases import make_engine, make_oracle, rel_l2

pytestmark = pytest.mark.gpu


def _peer_slabs(case, world, **kw):
    engs = [make_engine(case, rank=r, world=world, **kw) for r in range(world)]
    descs = [e.peer_export() for e in engs]
    py = case.periodic[1]
    for r, e in enumerate(engs):
        for side, q in ((0, r - 1), (1, r + 1)):
            if 0 <= q < world or py:
                e.peer_attach(side, descs[q % world])
    return engs


def test_c1_taylor_green_256_bgk_1000_steps_vs_oracle_and_analytic():
    """BASELINE configs[0]: engine vs oracle after the full 1000 steps, and the reference's metric within its calibrated band."""
    case = cases.Case("c1", 256, 256, cases.BGK, 1.0 / 6.0, (True, True), 0.04, "tg", scale=2)
    rho0, u0 = case.init_fields()
    o, e = make_oracle(case), make_engine(case)
    o.init(rho0, u0); e.init_fields(rho0, u0)
    o.step(1000); e.step(1000, macroscopics=True)
    (r_o, u_o), (r_e, u_e) = o.macroscopics(), e.macroscopics()
    s = e.taylor_green_error_sums(case.nu, 0.04 / 2, 1000.0)
    e.close()
    assert np.abs(r_e - r_o).max() <= 1e-5 and rel_l2(u_e, u_o) <= 1e-4, (np.abs(r_e - r_o).max(), rel_l2(u_e, u_o))
    err = 100 * np.sqrt(s[0] / s[1])
    assert 0.010 < err < 0.050, err            # 0.013-0.047 % over the run; reference CUDA on a B200: 0.0302 %


def test_c2_poiseuille_1024x256_mrt_vs_oracle():
    """BASELINE configs[1] at full size: body force, wet-node bounce-back walls, MRT with the reference's force-moment rows."""
    case = cases.Case("c2", 1024, 256, cases.MRT, 1.0 / 6.0, (True, False), 0.05, "pois", force=cases._pois_force(256))
    rho0, u0 = case.init_fields()
    o, e = make_oracle(case), make_engine(case)
    o.init(rho0, u0); e.init_fields(rho0, u0)
    o.step(400); e.step(400, macroscopics=True)
    (r_o, u_o), (r_e, u_e) = o.macroscopics(), e.macroscopics()
    f_o, f_e = o.populations(), e.populations()
    e.close()
    n = 400
    assert np.abs(f_e - f_o).max() <= 2e-6 * n ** 0.5 and np.abs(r_e - r_o).max() <= 1e-5 * n ** 0.5
    # The flow is driven by F = 1.0e-6 per step, i.e. each population of size 0.03 .. 0.44 receives an increment of only 15 .. 45 ulp
    # per step: in fp32 the rounding of that addition is a per-cent-level, evaluation-order dependent part of the acceleration
    # (the reference has the same limitation).  Engine and oracle evaluate the same formulas in different orders, so in this regime
    # they agree to 4.5e-4 per cell and 2.9e-4 in the x-averaged profile (measured) where other cases agree to 1e-5: the bar here is
    # 1e-3, with an absolute bar on the per-cell deviation.
    assert np.abs(u_e - u_o).max() <= 4e-8 * n ** 0.5 * 3, np.abs(u_e - u_o).max()
    assert rel_l2(u_e, u_o) <= 1e-3
    prof_e, prof_o = u_e[..., 0].astype(np.float64).mean(axis=1), u_o[..., 0].astype(np.float64).mean(axis=1)
    assert rel_l2(prof_e, prof_o) <= 1e-3, rel_l2(prof_e, prof_o)
    assert 2e-4 < prof_o[128] < 5e-4


@pytest.mark.parametrize("coll,steps", [(cases.CM, 40), (cases.CM_OPT, 12)])
def test_c3_lid_driven_4096_slab_invariance(coll, steps):
    """BASELINE configs[2] (4096^2 cavity, regularized BCs and corners): three peer-mapped slabs == one handle, bit for bit
    (OptimalAdapter: the grid sums are all-reduced in fp64, association differs)."""
    n = 4096
    nu = 0.1 * n / 1000.0
    case = cases.Case("c3", n, n, coll, nu, (False, False), 0.1, "lid")
    rho0, u0 = case.init_fields()
    one = make_engine(case)
    one.init_fields(rho0, u0)
    one.step(steps, macroscopics=True)
    r1, u1 = one.macroscopics()
    one.close()
    assert np.isfinite(r1).all() and np.isfinite(u1).all()
    engs = _peer_slabs(case, 3)
    for e in engs:
        e.init_fields(rho0, u0)
    for e in engs:
        e.sync()
    if coll == cases.CM_OPT:
        for i in range(steps):
            for e in engs:
                e.adapter_prepass()
            tot = sum(e.moment_sums() for e in engs)
            for e in engs:
                e.set_moment_sums(tot)
            for e in engs:
                e.step(1, macroscopics=(i == steps - 1))
    else:
        for e in engs:
            e.step(steps, macroscopics=True)
    for e in engs:
        e.sync()
    rs = np.concatenate([e.macroscopics()[0] for e in engs], axis=0)
    us = np.concatenate([e.macroscopics()[1] for e in engs], axis=0)
    for e in engs:
        e.close()
    tol = 0.0 if coll != cases.CM_OPT else 2e-7
    assert np.abs(rs - r1).max() <= tol and np.abs(us - u1).max() <= tol, (np.abs(rs - r1).max(), np.abs(us - u1).max())
    assert np.abs(u1[-1, n // 2, 0] - 0.1) < 0.02 and np.abs(u1).max() < 0.2        # the lid drives the top row at u_max


def test_c4_taylor_green_32768_analytic_and_mass():
    """BASELINE configs[3], one GPU: 2^30 cells, 64-bit offsets.  The analytic error is taken on the device (two doubles cross
    PCIe); with the stale-rest-population defect repaired (quirks = 0) mass is conserved to fp32 round-off."""
    import cuda_lbm_b200 as L
    n, scale, steps = 32768, 256.0, 60
    nu, u0 = np.float32(1.0 / 6.0), np.float32(0.04) / np.float32(scale)
    for quirks in (L.QK_REFERENCE, L.QK_FIXED):
        e = L.Engine(n, n, collision=L.BGK, viscosity=nu, periodic=(True, True), u_max=0.04, quirks=quirks)
        L._capi.check(L._capi.lib().lbm_reserve_macroscopics(e._h))
        e.init_taylor_green(nu, u0)
        s0 = e.taylor_green_error_sums(nu, u0, 0.0)
        m0 = e.total_mass()
        e.step(steps, macroscopics=True)
        s1 = e.taylor_green_error_sums(nu, u0, float(steps))
        m1 = e.total_mass()
        bpc = e.info().bytes_per_cell
        e.close()
        err0, err1 = 100 * np.sqrt(s0[0] / s0[1]), 100 * np.sqrt(s1[0] / s1[1])
        print(f"c4 quirks={quirks}: TG L2 error {err0:.5f} % at t=0, {err1:.5f} % after {steps} steps; mass drift {m1 / m0 - 1:.3e}; {bpc:.1f} B/cell")
        assert np.isfinite(s1).all() and abs(s1[1] / (0.5 * float(u0) ** 2 * n * n) - 1) < 1e-3      # sum |u_ref|^2 = N u0^2 / 2 (kx = ky, no decay yet)
        assert err0 < 1e-3 and err1 < 0.5
        assert abs(m0 / (float(n) * n) - 1) < 1e-6
        if quirks == L.QK_FIXED:
            assert abs(m1 / m0 - 1) < 1e-6, m1 / m0 - 1


def test_c5_cylinder_ibm_8192x2048_slab_invariance():
    """BASELINE configs[4] (Zou-He inlet, zero-gradient outflow, bounce-back walls, 804-marker cylinder, MRT): two peer-mapped
    slabs — the cylinder sits across their face — equal one handle bit for bit."""
    from oracle import oracle as O
    nx, ny, steps = 8192, 2048, 24
    case = cases.Case("c5", nx, ny, cases.MRT, cases._cyl_nu(ny), (False, False), 0.05, "cyl_ibm")
    cx, cy, r = case.cyl()
    case.bodies = [O.create_cylinder(cx, cy, r, 804)]
    rho0, u0 = case.init_fields()
    one = make_engine(case)
    one.init_fields(rho0, u0)
    one.step(steps, macroscopics=True)
    r1, u1 = one.macroscopics()
    info = one.info()
    one.close()
    assert info.num_markers == 804 and info.num_ibm_nodes > 804 and np.isfinite(u1).all()
    engs = _peer_slabs(case, 2)
    for e in engs:
        e.init_fields(rho0, u0)
    for e in engs:
        e.sync()
    for e in engs:
        e.step(steps, macroscopics=True)
    for e in engs:
        e.sync()
    rs = np.concatenate([e.macroscopics()[0] for e in engs], axis=0)
    us = np.concatenate([e.macroscopics()[1] for e in engs], axis=0)
    for e in engs:
        e.close()
    assert np.array_equal(rs, r1) and np.array_equal(us, u1), (np.abs(rs - r1).max(), np.abs(us - u1).max())
    assert abs(u1[ny // 2, 0, 0] - 0.05) < 5e-3            # Zou-He inlet column at u_max
