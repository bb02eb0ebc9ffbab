"""GPU: checkpoint / restart and the device-side validation reductions, through the C ABI.

Bars: a run continued from a checkpoint equals the uninterrupted run BIT FOR BIT (populations, rho, u) — at odd and even
AA phases, with boundary rings, IBM bodies, the OptimalAdapter means (both adapter modes) and peer-mapped slabs; the error
sums equal a float64 numpy evaluation of the same formula to 1e-6 relative (the device contracts dx*dx+dy*dy into an FMA and
uses its own cosf/sinf/expf for the analytic field).
"""
import os

import numpy as np
import pytest

import cases
from cases import make_engine, make_oracle

pytestmark = pytest.mark.gpu


def _fresh(case, adapter_mode=0, **kw):
    e = make_engine(case, adapter_mode=adapter_mode, **kw)
    return e


@pytest.mark.parametrize("name,adapter_mode,n1", [("g_tg_bgk", 0, 5), ("g_tg_bgk", 0, 4), ("g_pois_mrt", 0, 3), ("g_lid_cmopt", 0, 5),
                                                  ("g_lid_cmopt", 1, 5), ("g_lid_cmopt", 1, 6), ("g_cyl_ibm_mrt", 0, 7), ("g_cyl_flag_bgk", 0, 2)])
def test_restart_is_bit_identical(tmp_path, name, adapter_mode, n1):
    case = cases.BY_NAME[name]
    rho0, u0 = case.init_fields()
    n2 = 6
    path = tmp_path / "state.ckpt"
    a = _fresh(case, adapter_mode)
    a.init_fields(rho0, u0)
    a.step(n1)
    a.checkpoint_write(path)
    assert os.path.getsize(path) == a.checkpoint_bytes()
    a.step(n2, macroscopics=True)
    f_a, (r_a, u_a) = a.populations(), a.macroscopics()
    a.close()

    b = _fresh(case, adapter_mode)          # never initialised: everything comes from the file
    b.checkpoint_read(path)
    assert b.info().timestep == n1
    b.step(n2, macroscopics=True)
    f_b, (r_b, u_b) = b.populations(), b.macroscopics()
    b.close()
    assert np.isfinite(f_a).all()
    assert np.array_equal(f_a, f_b), np.abs(f_a - f_b).max()
    assert np.array_equal(r_a, r_b) and np.array_equal(u_a, u_b)


def test_restart_rejects_foreign_and_truncated_files(tmp_path):
    import cuda_lbm_b200 as L
    case = cases.BY_NAME["g_tg_bgk"]
    a = _fresh(case)
    a.init_fields(*case.init_fields())
    a.step(2)
    good = tmp_path / "good.ckpt"
    a.checkpoint_write(good)
    raw = open(good, "rb").read()
    (tmp_path / "short.ckpt").write_bytes(raw[: len(raw) // 2])
    (tmp_path / "junk.ckpt").write_bytes(b"not a checkpoint" * 64)
    for bad in ("short.ckpt", "junk.ckpt", "missing.ckpt"):
        with pytest.raises(L.LbmError):
            a.checkpoint_read(tmp_path / bad)
    other = cases.Case("other", 64, 24, cases.BGK, 1.0 / 6.0, (True, True), 0.04, "tg")
    o = _fresh(other)
    with pytest.raises(L.LbmError) as ei:
        o.checkpoint_read(good)
    assert "different grid" in str(ei.value)
    o.close()
    # the handle that failed to read a truncated file is still usable after a good read
    a.checkpoint_read(good)
    a.step(1)
    a.sync()
    a.close()


def test_peer_mapped_slabs_restart(tmp_path):
    """Two peer-mapped slabs written at an odd step and continued in fresh handles equal the uninterrupted pair bit for bit."""
    case = cases.BY_NAME["g_tg_mrt"]
    rho0, u0 = case.init_fields()

    def pair():
        engs = [make_engine(case, rank=r, world=2) for r in range(2)]
        d = [e.peer_export() for e in engs]
        for r, e in enumerate(engs):
            e.peer_attach(0, d[1 - r]); e.peer_attach(1, d[1 - r])
        return engs

    def run(engs, n, macros=False):
        for e in engs:
            e.step(n, macroscopics=macros)
        for e in engs:
            e.sync()

    a = pair()
    for e in a:
        e.init_fields(rho0, u0)
    for e in a:
        e.sync()
    run(a, 3)
    for r, e in enumerate(a):
        e.checkpoint_write(tmp_path / f"slab{r}.ckpt")
    run(a, 4, True)
    f_a = np.concatenate([e.populations() for e in a], axis=0)
    for e in a:
        e.close()
    b = pair()
    for r, e in enumerate(b):
        e.checkpoint_read(tmp_path / f"slab{r}.ckpt")
    run(b, 4, True)
    f_b = np.concatenate([e.populations() for e in b], axis=0)
    for e in b:
        e.close()
    assert np.array_equal(f_a, f_b), np.abs(f_a - f_b).max()


def _tg_analytic(nx, ny, nu, u0, t):
    """TaylorGreenValidation (reference taylorGreenFunctors.cuh:66-81) in float64."""
    y, x = np.meshgrid(np.arange(ny) + 0.5, np.arange(nx) + 0.5, indexing="ij")
    kx, ky = 2 * np.pi / nx, 2 * np.pi / ny
    decay = np.exp(-t * nu * (kx * kx + ky * ky))
    ux = -u0 * np.sqrt(ky / kx) * np.cos(kx * x) * np.sin(ky * y) * decay
    uy = u0 * np.sqrt(kx / ky) * np.sin(kx * x) * np.cos(ky * y) * decay
    return np.stack([ux, uy], axis=-1)


@pytest.mark.parametrize("world", [1, 3])
def test_error_sums_match_numpy(world):
    import torch
    case = cases.Case("tg_val", 96, 60, cases.BGK, 1.0 / 6.0, (True, True), 0.04, "tg")
    rho0, u0 = case.init_fields()
    nsteps = 30
    ref64 = _tg_analytic(case.nx, case.ny, float(case.nu), 0.04, float(nsteps))
    ref32 = ref64.astype(np.float32)
    engs = [make_engine(case, rank=r, world=world) for r in range(world)]
    if world > 1:
        d = [e.peer_export() for e in engs]
        for r, e in enumerate(engs):
            e.peer_attach(0, d[(r - 1) % world]); e.peer_attach(1, d[(r + 1) % world])
    for e in engs:
        e.init_fields(rho0, u0)
    for e in engs:
        e.sync()
    for e in engs:
        e.step(nsteps, macroscopics=True)
    for e in engs:
        e.sync()
    tot_tg, tot_fld = np.zeros(2), np.zeros(2)
    u_all, means = [], []
    for e in engs:
        sl = slice(e.y0, e.y0 + e.ny_local)
        d_ref = torch.from_numpy(np.ascontiguousarray(ref32[sl])).cuda()
        tot_fld += e.velocity_error_sums(d_ref.data_ptr())
        tot_tg += e.taylor_green_error_sums(case.nu, 0.04, nsteps)
        u_all.append(e.macroscopics()[1])
        means.append(e.row_mean_velocity())
    for e in engs:
        e.close()
    u = np.concatenate(u_all, axis=0).astype(np.float64)
    want_fld = np.array([((u - ref32.astype(np.float64)) ** 2).sum(), (ref32.astype(np.float64) ** 2).sum()])
    want_tg = np.array([((u - ref64) ** 2).sum(), (ref64 ** 2).sum()])
    assert np.allclose(tot_fld, want_fld, rtol=1e-6, atol=0), (tot_fld, want_fld)
    assert np.allclose(tot_tg, want_tg, rtol=2e-4, atol=0), (tot_tg, want_tg)     # err sum: cancellation amplifies the 1e-7 of cosf/expf
    assert abs(tot_tg[1] / want_tg[1] - 1) < 1e-6
    # the reference's metric (taylorGreenScenario.cuh:87) from the device sums vs from the host fields
    assert abs(100 * np.sqrt(tot_tg[0] / tot_tg[1]) - 100 * np.sqrt(want_tg[0] / want_tg[1])) < 1e-4
    mx = np.concatenate([m[0] for m in means]); my = np.concatenate([m[1] for m in means])
    assert np.allclose(mx, u[..., 0].mean(axis=1), rtol=0, atol=1e-12) and np.allclose(my, u[..., 1].mean(axis=1), rtol=0, atol=1e-12)


def test_validation_needs_current_macroscopics():
    import cuda_lbm_b200 as L
    case = cases.BY_NAME["g_tg_bgk"]
    e = make_engine(case)
    e.init_fields(*case.init_fields())
    e.step(2)
    with pytest.raises(L.LbmError):
        e.taylor_green_error_sums(case.nu, 0.04, 2.0)
    with pytest.raises(L.LbmError):
        e.row_mean_velocity()
    e.close()


def test_poiseuille_profile_metric_from_row_means():
    """PoiseuilleScenario::compute_error (poiseuilleScenario.cuh:55-77) from lbm_row_mean_velocity equals the host evaluation
    over the full velocity field, and both equal the oracle's."""
    case = cases.Case("pois_val", 64, 32, cases.MRT, 1.0 / 6.0, (True, False), 0.05, "pois", force=cases._pois_force(32))
    rho0, u0 = case.init_fields()
    e, o = make_engine(case), make_oracle(case)
    e.init_fields(rho0, u0); o.init(rho0, u0)
    e.step(200, macroscopics=True); o.step(200)
    mx, _ = e.row_mean_velocity()
    u_e, u_o = e.macroscopics()[1], o.macroscopics()[1]
    e.close()
    y = np.arange(case.ny, dtype=np.float64)
    prof = (8.0 * (1.0 / 6.0) * 0.05 / case.ny ** 2) / (2.0 / 6.0) * y * (case.ny - y)

    def metric(mean_ux):
        return np.sqrt(((mean_ux - prof) ** 2).sum() / case.ny) * 100.0 / 0.05

    m_dev, m_host, m_orc = metric(mx), metric(u_e[..., 0].astype(np.float64).mean(axis=1)), metric(u_o[..., 0].astype(np.float64).mean(axis=1))
    assert abs(m_dev - m_host) < 1e-9
    assert abs(m_dev - m_orc) < 1e-3 * m_orc
