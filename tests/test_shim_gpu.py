"""GPU: main.cu-style drivers built over the header shim (examples/_bin/, built by `make -C examples`).

 * ref_* binaries are the reference's OWN scenario files (taylorGreen, poiseuille, lidDrivenCavity — unmodified) compiled against
   include/cuda-lbm/ and run on the B200 engine; they are compared with the reference's own CUDA solver built from the same
   scenario constants (oracle/_ref/bin/, oracle/build_ref.sh) after N steps:  max |rho - rho_ref| and rel-L2(u) in fp32.
   This is BASELINE.json's parity statement ("results must match the reference's own CUDA solver on identical scenarios").
 * ex_* binaries are this repo's scenario files for the BASELINE configurations: analytic checks (Taylor-Green decay,
   Poiseuille profile), the per-step driver protocol against LBM::run, the VTK / raw writers.

Tolerances (fp32, stated per assertion): the engine evaluates the reference's formulas with different association / FMA
contraction, so after N steps fields differ by accumulated round-off, not bit for bit.
"""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, "examples", "_bin")
REF = os.path.join(ROOT, "oracle", "_ref", "bin")


def need(path):
    if not os.path.exists(path):
        pytest.skip(f"{os.path.relpath(path, ROOT)} not built (make -C examples / oracle/build_ref.sh in the build container)")
    return path


def run_shim(name, cwd, *args):
    r = subprocess.run([need(os.path.join(EX, name))] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    m = re.search(r"SHIM_RESULT (.*)", r.stdout)
    assert m, r.stdout[-2000:]
    res = dict(kv.split("=") for kv in m.group(1).split())
    errors = [(int(t), float(e)) for t, e in re.findall(r"\[(\d+)\]: error, ([-0-9.eE+naif]+)%", r.stdout)]
    return res, errors, r.stdout


def shim_fields(cwd, t, nx, ny):
    rho = np.fromfile(os.path.join(cwd, "output", "density", f"density_{t}.bin"), np.float32).reshape(ny, nx)
    u = np.fromfile(os.path.join(cwd, "output", "velocity", f"velocity_{t}.bin"), np.float32).reshape(ny, nx, 2)
    return rho, u


def ref_fields(name, cwd, steps, nx, ny):
    r = subprocess.run([need(os.path.join(REF, name)), str(steps), str(cwd), "ref", str(steps)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rho = np.fromfile(os.path.join(cwd, f"ref_t{steps}.rho.bin"), np.float32).reshape(ny, nx)
    u = np.fromfile(os.path.join(cwd, f"ref_t{steps}.u.bin"), np.float32).reshape(ny, nx, 2)
    return rho, u


def rel_l2(a, b):
    return float(np.sqrt(((a.astype(np.float64) - b) ** 2).sum() / max((b.astype(np.float64) ** 2).sum(), 1e-300)))


# (shim binary, reference CUDA binary, nx, ny, steps, tol max|d rho|, tol rel-L2 u)
# measured on the B200 (profiles/r01_shim_parity.txt): TG 6.0e-5, the others below 1e-5
PAIRS = [("ref_tg_256", "c1_tg_bgk_256", 256, 256, 1000, 5e-6, 1e-4),
         ("ref_lid_129", "s_lid_bgk_129", 129, 129, 1000, 2e-5, 1e-4),
         ("ref_pois_150x100", "s_pois_bgk_150x100", 150, 100, 500, 2e-5, 2e-4),
         # the reference's flowPastCylinderScenario.cuh (two 16-marker IBM cylinders, Zou-He inlet, zero-gradient outflow) after the
         # one-token build fix it needs in the reference too (`BGK` -> `BGK<2>`)
         ("ref_cyl_256x128", "s_cyl_bgk_256x128", 256, 128, 500, 2e-5, 2e-4)]


@pytest.mark.parametrize("shim,ref,nx,ny,steps,tol_rho,tol_u", PAIRS, ids=[p[0] for p in PAIRS])
def test_reference_scenarios_on_the_engine_match_the_reference_cuda_solver(tmp_path, shim, ref, nx, ny, steps, tol_rho, tol_u):
    res, _, _ = run_shim(shim, tmp_path, "--steps", steps, "--save-int", steps, "--dump")
    rho_s, u_s = shim_fields(tmp_path, steps, nx, ny)
    rho_r, u_r = ref_fields(ref, tmp_path, steps, nx, ny)
    assert np.isfinite(rho_s).all() and np.isfinite(u_s).all()
    d_rho, d_u = float(np.abs(rho_s - rho_r).max()), rel_l2(u_s, u_r)
    print(f"{shim} vs {ref} after {steps} steps: max|drho|={d_rho:.3e} relL2(u)={d_u:.3e} max|du|={float(np.abs(u_s - u_r).max()):.3e}")
    assert d_rho <= tol_rho, d_rho
    assert d_u <= tol_u, d_u


def tg_analytic_error_pct(u, t, nx, ny, scale, nu=1.0 / 6.0, u_max=0.04):
    y, x = np.meshgrid(np.arange(ny) + 0.5, np.arange(nx) + 0.5, indexing="ij")
    kx, ky = 2 * np.pi / nx, 2 * np.pi / ny
    dec = np.exp(-t * nu * (kx * kx + ky * ky))
    u0 = u_max / scale
    ax = -u0 * np.sqrt(ky / kx) * np.cos(kx * x) * np.sin(ky * y) * dec
    ay = u0 * np.sqrt(kx / ky) * np.sin(kx * x) * np.cos(ky * y) * dec
    err = ((u[..., 0] - ax) ** 2 + (u[..., 1] - ay) ** 2).sum()
    return float(np.sqrt(err / (ax ** 2 + ay ** 2).sum()) * 100)


def test_taylor_green_256_decay_no_worse_than_the_reference(tmp_path):
    """BASELINE config 1: 256 x 256, BGK, 1000 steps, L2 error against the analytic decay every 100 steps."""
    res, errors, _ = run_shim("ex_c1_tg_256", tmp_path, "--steps", 1000, "--save-int", 100, "--dump")
    assert [t for t, _ in errors] == list(range(100, 1001, 100))
    assert all(0.0 < e < 0.06 for _, e in errors), errors          # SURVEY.md 8c calibration: 0.013 .. 0.047 %
    rho, u = shim_fields(tmp_path, 1000, 256, 256)
    mine = tg_analytic_error_pct(u, 1000, 256, 256, 2)
    assert abs(mine - errors[-1][1]) < 2e-3                         # the scenario's own metric agrees with the numpy one
    _, u_r = ref_fields("c1_tg_bgk_256", tmp_path, 1000, 256, 256)
    ref = tg_analytic_error_pct(u_r, 1000, 256, 256, 2)
    print(f"TG 256^2 analytic L2 error after 1000 steps: engine {mine:.5f} %, reference CUDA {ref:.5f} %")
    assert mine <= ref * 1.10 + 1e-4
    assert abs(float(res["mass_per_node"]) - 1.0) < 1e-4


@pytest.mark.parametrize("name", ["ex_tg_mrt_256", "ex_tg_cm_256"])
def test_taylor_green_other_operators_track_the_analytic_decay(tmp_path, name):
    res, errors, _ = run_shim(name, tmp_path, "--steps", 400, "--save-int", 200)
    assert all(0.0 < e < 0.1 for _, e in errors), errors


def test_optimal_adapter_matches_the_reference_including_its_instability(tmp_path):
    """CM<2,OptimalAdapter> relaxes the three highest central moments at 1/(3 tau* + 1/2) ~ 1.9 .. 1.98 (SURVEY.md Appendix
    A-D10).  On a Taylor-Green box (64^2 here; 256^2 behaves the same, profiles/r01_shim_parity.txt) the reference's own CUDA solver leaves the physical branch within ~100 steps and
    overflows; the engine reproduces both phases: close agreement while the reference is finite, non-finite afterwards."""
    res, errors, _ = run_shim("ex_tg_cmopt_64", tmp_path, "--steps", 30, "--save-int", 30, "--dump")
    rho_s, u_s = shim_fields(tmp_path, 30, 64, 64)
    rho_r, u_r = ref_fields("s_tg_cmopt_64", tmp_path, 30, 64, 64)
    assert np.isfinite(rho_r).all() and np.isfinite(rho_s).all()
    d_rho, d_u = float(np.abs(rho_s - rho_r).max()), rel_l2(u_s, u_r)
    print(f"TG 64^2 CM<OptimalAdapter> after 30 steps vs reference CUDA: max|drho|={d_rho:.3e} relL2(u)={d_u:.3e}")
    assert d_rho < 1e-4 and d_u < 2e-3          # own init functor (expf/cosf association differs from the reference's double-promoted form)
    # The reference's debug build prints from device code once values exceed its VALUE_THRESHOLD (lbm.cuh:21), which makes it crawl after
    # the blow-up: small grid, and stop shortly after the point where the oracle shows the overflow (~70 steps at 64^2).
    n = 80
    rho_r, _ = ref_fields("s_tg_cmopt_64", tmp_path, n, 64, 64)
    run_shim("ex_tg_cmopt_64", tmp_path, "--steps", n, "--save-int", n, "--dump")
    rho_s, _ = shim_fields(tmp_path, n, 64, 64)
    off = lambda r: (not np.isfinite(r).all()) or float(np.abs(r - 1).max()) > 0.5      # noqa: E731
    print(f"after {n} steps: reference left the physical branch: {off(rho_r)}, engine: {off(rho_s)}")
    assert off(rho_r), "the reference became stable: revisit DESIGN.md's note on OptimalAdapter"
    assert off(rho_s)

def test_driver_protocol_equals_run(tmp_path):
    """The ten per-step calls of src/main.cu:96-114 and LBM::run(n) enqueue the same launches: identical results."""
    a, _, _ = run_shim("ex_tg_mrt_256", tmp_path, "--steps", 37, "--save-int", 10)
    b, _, _ = run_shim("ex_tg_mrt_256", tmp_path, "--steps", 37, "--save-int", 10, "--fast")
    # 37 steps in one LBM::run call: two replays of the 16-step CUDA graph (captured on the handle's own stream, bridged to the
    # legacy default stream the shim works on) + 5 ordinary launches
    c, _, _ = run_shim("ex_tg_mrt_256", tmp_path, "--steps", 37, "--save-int", 37, "--fast")
    for k in ("error_pct", "mass_per_node", "mean_rho", "sum_u2"):
        assert a[k] == b[k] == c[k], (k, a[k], b[k], c[k])


def test_checkpoint_restart_through_the_driver(tmp_path):
    """main --save-ckpt after 600 steps, then --load-ckpt + 400 steps in a new process: the same result line as 1000 steps in
    one go (LBM::save_checkpoint / load_checkpoint<S> over lbm_checkpoint_write / lbm_checkpoint_read)."""
    ck = str(tmp_path / "state.ckpt")
    run_shim("ex_tg_mrt_256", tmp_path, "--steps", 600, "--save-int", 600, "--fast", "--save-ckpt", ck)
    b, errors, out = run_shim("ex_tg_mrt_256", tmp_path, "--steps", 400, "--save-int", 400, "--fast", "--load-ckpt", ck)
    a, _, _ = run_shim("ex_tg_mrt_256", tmp_path, "--steps", 1000, "--save-int", 1000, "--fast")
    assert "restarted from" in out and "at step 600" in out
    for k in ("error_pct", "mass_per_node", "mean_rho", "sum_u2"):
        assert a[k] == b[k], (k, a[k], b[k])


def test_poiseuille_profile(tmp_path):
    """64 x 32 channel, MRT, body force, wet-node bounce-back walls, to steady state (H^2/nu = 6.1e3 steps).
    The reference's metric assumes walls at y = 0 and y = NY (7.0 % predicted at NY = 32); against the
    profile for walls on the first / last node row the same field is within 1 %."""
    res, errors, out = run_shim("ex_pois_64x32", tmp_path, "--steps", 40000, "--save-int", 40000, "--fast", "--dump")
    assert 6.0 < errors[-1][1] < 7.6, errors
    rho, u = shim_fields(tmp_path, 40000, 64, 32)
    ny = 32
    y = np.arange(ny)
    prof = u[..., 0].mean(axis=1)
    wet = (8 * (1 / 6) * 0.05 / ny ** 2) / (2 * (1 / 6)) * y * (ny - 1 - y)
    err_wet = np.sqrt(((prof - wet) ** 2).mean()) * 100 / 0.05
    print(f"Poiseuille 64x32: reference metric {errors[-1][1]:.3f} %, wet-node profile error {err_wet:.3f} %")
    assert err_wet < 1.5
    assert np.abs(u[..., 0] - prof[:, None]).max() < 5e-4          # x-invariant up to the corner-node defect (Appendix A-D11)


def test_cavity_and_cylinder_run(tmp_path):
    res, _, _ = run_shim("ex_lid_129", tmp_path, "--steps", 2000, "--save-int", 1000)
    assert np.isfinite(float(res["sum_u2"])) and float(res["sum_u2"]) > 0 and abs(float(res["mean_rho"]) - 1) < 0.05
    res, _, out = run_shim("ex_cyl_256x128", tmp_path, "--steps", 1000, "--save-int", 500)
    assert np.isfinite(float(res["sum_u2"])) and float(res["sum_u2"]) > 0


def test_vtk_and_raw_writers(tmp_path):
    """save_vtk writes the reference's .vti layout (lbm.cuh:262-343): UInt64 block length + Float32 Density, then
    UInt64 + Float32 x3 Velocity; save_macroscopics writes density_<t>.bin / velocity_<t>.bin (lbm.cuh:173-204)."""
    nx, ny, t = 64, 32, 20
    run_shim("ex_pois_64x32", tmp_path, "--steps", t, "--save-int", t, "--vtk", "--dump")
    rho, u = shim_fields(tmp_path, t, nx, ny)
    raw = open(os.path.join(tmp_path, "output", "vtk", f"sim_data_{t:06d}.vti"), "rb").read()
    head, _, rest = raw.partition(b"<AppendedData encoding=\"raw\">\n   _")
    assert b'WholeExtent="0 63 0 31 0 0"' in head and b'Name="Density" format="appended" offset="0"' in head
    assert f'NumberOfComponents="3" format="appended" offset="{nx * ny * 4}"'.encode() in head
    (n0,) = struct.unpack("<Q", rest[:8])
    assert n0 == nx * ny * 4
    d = np.frombuffer(rest[8:8 + n0], np.float32).reshape(ny, nx)
    (n1,) = struct.unpack("<Q", rest[8 + n0:16 + n0])
    assert n1 == nx * ny * 12
    v = np.frombuffer(rest[16 + n0:16 + n0 + n1], np.float32).reshape(ny, nx, 3)
    assert np.array_equal(d, rho) and np.array_equal(v[..., :2], u) and not v[..., 2].any()
    assert rest[16 + n0 + n1:].strip().endswith(b"</VTKFile>")


def test_call_orders_and_per_step_forces(tmp_path):
    """examples/order_check.cu: (1) save_checkpoint() then save_vtk(), and update_macroscopics() after the next step's increase_ts(),
    are legal as in the reference (its d_rho / d_u are always current); (2) a scenario that declares time_dependent_forces gets
    Init::apply_forces re-evaluated every step, bit-identical to setting the force anew through the C ABI."""
    r = subprocess.run([need(os.path.join(EX, "t_order_check")), str(tmp_path / "oc.ckpt")], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    print(r.stdout[-1500:])
    assert r.returncode == 0 and "ORDER_CHECK PASSED" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def _run_slabs_env(name, cwd, gpus, *args):
    env = dict(os.environ, LBM_B200_GPUS=str(gpus))
    cmd = [need(os.path.join(EX, name))] + [str(a) for a in args]
    r = subprocess.run(cmd, cwd=cwd, env=env, capture_output=True, text=True, timeout=900)
    if r.returncode != 0 and gpus > 1 and "handshake timed out" in r.stdout + r.stderr:
        # Several slabs that SHARE one device (this box; a test topology — in production every slab has its own GPU) spin for one another
        # on that device, and about one such run in fifty ends in the engine's 10 s handshake time-out instead of a result (2 of 96 runs of
        # ex_cyl_256x128 on 4 slabs, profiles/r02_shared_device_timeouts.txt).  The time-out is reported, never a wrong result: run once more.
        print("NOTE: handshake time-out with slabs sharing one device, running once more:", (r.stdout + r.stderr)[-300:])
        r = subprocess.run(cmd, cwd=cwd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    m = re.search(r"SHIM_RESULT (.*)", r.stdout)
    assert m, r.stdout[-2000:]
    return dict(kv.split("=") for kv in m.group(1).split()), r.stdout


# (binary, nx, ny, steps, slabs, exact) — the cylinder of ex_cyl_256x128 sits on the face between slabs 1 and 2 of 4
MULTI = [("ex_tg_mrt_256", 256, 256, 37, 3, True), ("ex_cyl_256x128", 256, 128, 60, 4, True), ("ex_pois_64x32", 64, 32, 50, 2, True),
         ("ex_tg_cmopt_64", 64, 64, 20, 4, False), ("ex_lid_129", 129, 129, 40, 5, True)]


@pytest.mark.parametrize("name,nx,ny,steps,slabs,exact", MULTI, ids=[m[0] for m in MULTI])
@pytest.mark.parametrize("fast", [False, True], ids=["protocol", "run"])
def test_scenario_drivers_on_several_slabs_equal_one_slab(tmp_path, name, nx, ny, steps, slabs, exact, fast):
    """LBM_B200_GPUS=N: the same unchanged main.cu-style driver runs the scenario on N y-slabs (one engine handle, host thread and
    stream per slab, all slabs peer-mapped; on this box they share the GPU, on an 8-GPU box each gets its own).  Fields after
    `steps` steps equal the one-slab run bit for bit (OptimalAdapter: the grid sums are added in another order, 2e-7)."""
    extra = ("--fast",) if fast else ()
    d1, dn = tmp_path / "one", tmp_path / "many"
    d1.mkdir(); dn.mkdir()
    a, _ = _run_slabs_env(name, d1, 1, "--steps", steps, "--save-int", steps, "--dump", *extra)
    b, out = _run_slabs_env(name, dn, slabs, "--steps", steps, "--save-int", steps, "--dump", *extra)
    assert b["gpus"] == str(slabs) and f"{slabs} y-slabs" in out
    rho1, u1 = shim_fields(d1, steps, nx, ny)
    rhon, un = shim_fields(dn, steps, nx, ny)
    assert np.isfinite(rho1).all()
    tol = 0.0 if exact else 2e-7
    assert float(np.abs(rhon - rho1).max()) <= tol and float(np.abs(un - u1).max()) <= tol, (float(np.abs(rhon - rho1).max()), float(np.abs(un - u1).max()))
    if exact:
        assert a["error_pct"] == b["error_pct"] and a["sum_u2"] == b["sum_u2"]


def test_checkpoint_restart_on_several_slabs(tmp_path):
    ck = str(tmp_path / "state.ckpt")
    _run_slabs_env("ex_tg_mrt_256", tmp_path, 3, "--steps", 30, "--save-int", 30, "--fast", "--save-ckpt", ck)
    assert all(os.path.exists(f"{ck}.slab{g}") for g in range(3))
    b, out = _run_slabs_env("ex_tg_mrt_256", tmp_path, 3, "--steps", 20, "--save-int", 20, "--fast", "--load-ckpt", ck)
    a, _ = _run_slabs_env("ex_tg_mrt_256", tmp_path, 1, "--steps", 50, "--save-int", 50, "--fast")
    assert "at step 30" in out
    for k in ("error_pct", "mean_rho", "sum_u2"):
        assert a[k] == b[k], (k, a[k], b[k])


def test_ghia_validation_of_the_converged_cavity(tmp_path):
    """SURVEY.md 8f-1 / lidDrivenCavityScenario.cuh:88-157: the reference validates its cavity against the centre-line tables of
    Ghia, Ghia & Shin (1982).  (a) The reference's OWN scenario file (129^2, Re = 100, BGK, regularized walls) on the engine, run to
    convergence: its own host-side metric and the same metric with the 2 x 17 samples gathered on the device (lbm_sample_velocity).
    (b) Re = 1000 with CM<2,NoAdapter> (the reference's functors and tables, constants chosen by -D).  Calibration with the CPU
    oracle (same arithmetic): 1.32 % at Re = 100 from 60 000 steps on, 1.55 % at Re = 1000 after 120 000."""
    res, errors, out = run_shim("ref_lid_129", tmp_path, "--steps", 60000, "--save-int", 60000, "--fast")
    host = errors[-1][1]
    m = re.search(r"samples gathered on the device, ([0-9.]+)%", out)
    assert m, out[-1500:]
    dev = float(m.group(1))
    print(f"Ghia Re=100 BGK 129^2 after 60000 steps: reference metric {host:.4f} % (host), {dev:.4f} % (device samples)")
    assert 1.0 < host < 1.7 and abs(host - dev) < 2e-3, (host, dev)
    res, errors, out = run_shim("ref_ghia_re1000_cm_129", tmp_path, "--steps", 120000, "--save-int", 120000, "--fast")
    print(f"Ghia Re=1000 CM 129^2 after 120000 steps: {errors[-1][1]:.4f} %")
    assert 1.0 < errors[-1][1] < 2.0, errors
