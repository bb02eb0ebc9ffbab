"""Parity cases shared by the oracle-vs-golden tests (CPU) and the engine-vs-oracle/golden tests (GPU).

Every case mirrors one line of oracle/ref_cuda/configs.txt (the reference CUDA build that produced
tests/golden/<name>.npz on a B200) and the Scenario struct oracle/ref_cuda/ref_driver.cu builds for it.
Constants are computed in float32 exactly as the reference's `static constexpr float` members are.
"""
import numpy as np

f32 = np.float32
BGK, MRT, CM, CM_OPT = 0, 1, 2, 3
FLUID, BOUNCE_BACK, ZOU_HE_TOP, ZOU_HE_LEFT, CYLINDER, ZG_OUTFLOW, PRESSURE_OUTLET = 0, 1, 2, 3, 6, 7, 8
REG_INLET_TOP, REG_BB, REG_BB_CORNER = 9, 11, 12


def ref_S(coll, omega):
    om = f32(omega)
    if coll >= CM:
        return np.array([0, 0, 0, 1, om, om, 1, 1, 1], f32)
    return np.array([0, om, om, 0, om, 0, om, om, om], f32)


def omega_of(nu):
    return f32(1.0) / (f32(3) * f32(nu) + f32(0.5))


class Case:
    def __init__(self, name, nx, ny, coll, nu, periodic, u_max, kind, S=None, force=(0.0, 0.0), np_markers=0, scale=1,
                 steps_f=(0, 1, 2, 3), steps_m=(1, 2, 3, 10, 100)):
        self.name, self.nx, self.ny, self.coll, self.nu = name, nx, ny, coll, f32(nu)
        self.periodic, self.u_max, self.kind = periodic, f32(u_max), kind
        self.omega = omega_of(nu)
        self.S = ref_S(coll, self.omega) if S is None else np.asarray(S, f32)
        self.force = force
        self.np_markers = np_markers
        self.scale = scale
        self.steps_f, self.steps_m = steps_f, steps_m

    # Boundary functor of the case evaluated on the grid (int32 [ny, nx])
    def flags(self):
        nx, ny = self.nx, self.ny
        y, x = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
        f = np.zeros((ny, nx), np.int32)
        k = self.kind
        if k == "tg":
            pass
        elif k == "pois":
            f[(y == 0) | (y == ny - 1)] = BOUNCE_BACK
        elif k == "lid":
            f[(x == 0) | (x == nx - 1) | (y == 0)] = REG_BB
            f[y == ny - 1] = REG_INLET_TOP
            f[((x == 0) | (x == nx - 1)) & ((y == 0) | (y == ny - 1))] = REG_BB_CORNER
        elif k == "lidzh":
            f[(x == 0) | (x == nx - 1) | (y == 0)] = BOUNCE_BACK
            f[y == ny - 1] = ZOU_HE_TOP
            f[((x == 0) | (x == nx - 1)) & ((y == 0) | (y == ny - 1))] = BOUNCE_BACK
        elif k in ("cyl_ibm", "cyl_flag"):
            if k == "cyl_flag":
                cx, cy, r = self.cyl()
                dx, dy = x.astype(f32) - cx, y.astype(f32) - cy
                f[(dx * dx + dy * dy) <= (r * r)] = CYLINDER
                f[x == nx - 1] = PRESSURE_OUTLET
            else:
                f[x == nx - 1] = ZG_OUTFLOW
            f[x == 0] = ZOU_HE_LEFT
            f[(y == 0) | (y == ny - 1)] = BOUNCE_BACK
        return f

    def cyl(self):
        D = f32(self.ny / 8.0)
        return f32(3.0) * D, f32(self.ny / 2.0), D / f32(2.0)

    def markers(self):
        """IBMBody::points of the case's bodies: one [n,2] array, or a list of them (one lbm_add_body call each)."""
        if getattr(self, "bodies", None) is not None:
            return self.bodies
        if self.kind != "cyl_ibm":
            return None
        from oracle import oracle as O
        cx, cy, r = self.cyl()
        return O.create_cylinder(cx, cy, r, self.np_markers)

    def init_fields(self):
        """rho,u as the case's Init functor writes them (CPU evaluation; golden tests start from the dumped t=0 state instead)."""
        if self.kind == "tg":
            from oracle import oracle as O
            return O.taylor_green_init(self.nx, self.ny, self.nu, f32(self.u_max) / f32(self.scale))
        return np.ones((self.ny, self.nx), f32), np.zeros((self.ny, self.nx, 2), f32)


def _cyl_nu(ny):
    return f32(0.05) * f32(ny / 8.0) / f32(50.0)


def _pois_force(ny):
    return (float(f32(8.0) * f32(1.0 / 6.0) * f32(0.05) / f32(ny * ny)), 0.0)


_tg_om = omega_of(1.0 / 6.0)
CASES = [
    Case("g_tg_bgk", 32, 24, BGK, 1.0 / 6.0, (True, True), 0.04, "tg"),
    Case("g_tg_mrt", 32, 24, MRT, 1.0 / 6.0, (True, True), 0.04, "tg", S=[0, 1.0, 1.4, 0, 1.2, 0, 1.9, _tg_om, _tg_om]),
    Case("g_tg_cm", 32, 24, CM, 1.0 / 6.0, (True, True), 0.04, "tg"),
    Case("g_tg_cmopt", 32, 24, CM_OPT, 1.0 / 6.0, (True, True), 0.04, "tg", steps_m=(1, 2, 3, 10, 30)),
    Case("g_pois_bgk", 32, 16, BGK, 1.0 / 6.0, (True, False), 0.05, "pois", force=_pois_force(16)),
    Case("g_pois_mrt", 32, 16, MRT, 1.0 / 6.0, (True, False), 0.05, "pois", force=_pois_force(16)),
    Case("g_pois_cm", 32, 16, CM, 1.0 / 6.0, (True, False), 0.05, "pois", force=_pois_force(16)),
    Case("g_lid_bgk", 33, 33, BGK, 0.03, (False, False), 0.1, "lid"),
    Case("g_lid_cm", 33, 33, CM, 0.03, (False, False), 0.1, "lid"),
    Case("g_lid_cmopt", 33, 33, CM_OPT, 0.03, (False, False), 0.1, "lid", steps_m=(1, 2, 3, 10, 30)),
    Case("g_lidzh_mrt", 33, 33, MRT, 0.03, (False, False), 0.1, "lidzh"),
    Case("g_cyl_ibm_mrt", 96, 48, MRT, _cyl_nu(48), (False, False), 0.05, "cyl_ibm", np_markers=16, steps_f=(0, 1, 2)),
    Case("g_cyl_ibm_bgk", 96, 48, BGK, _cyl_nu(48), (False, False), 0.05, "cyl_ibm", np_markers=40, steps_f=(0, 1, 2)),
    Case("g_cyl_flag_bgk", 96, 48, BGK, _cyl_nu(48), (False, False), 0.05, "cyl_flag", steps_f=(0, 1, 2)),
]
BY_NAME = {c.name: c for c in CASES}


def make_oracle(case, quirks=63):
    from oracle import oracle as O
    o = O.Oracle(case.nx, case.ny, coll=case.coll, viscosity=case.nu, S=case.S, periodic=case.periodic, u_max=case.u_max,
                 force=case.force, quirks=quirks)
    o.set_flags(case.flags())
    m = case.markers()
    if m is not None:
        o.set_markers(np.concatenate(m, axis=0) if isinstance(m, list) else m)     # the reference keeps one marker array for all bodies (IBMManager.cuh:54-109)
    return o


def make_engine(case, quirks=63, adapter_mode=0, **kw):
    import cuda_lbm_b200 as L
    e = L.Engine(case.nx, case.ny, collision=case.coll, viscosity=case.nu, S=case.S, periodic=case.periodic,
                 u_max=case.u_max, force=case.force, quirks=quirks, adapter_mode=adapter_mode, **kw)
    e.set_flags(case.flags())
    m = case.markers()
    if m is not None:
        for body in (m if isinstance(m, list) else [m]):
            e.add_body(body)
    return e


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = np.sqrt(np.sum((a - b) ** 2))
    n = np.sqrt(np.sum(b ** 2))
    return d / n if n > 0 else d
