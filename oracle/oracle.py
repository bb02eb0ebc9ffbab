"""TEST INFRASTRUCTURE — ctypes binding of oracle/liblbm_oracle.so (the CPU restatement of the
reference's D2Q9 path, see lbm_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs may import this module; the product package
(cuda_lbm_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# collision operators (order used everywhere in this repo)
BGK, MRT, CM, CM_OPT = 0, 1, 2, 3
# quirk bits, lbm_oracle.c
QK_D1_STALE_F0, QK_D2_MRT_ROWS, QK_D3_ZOUHE_RHO, QK_D7_IBM_CLIP, QK_D8_IBM_2X2, QK_D11_BB_RAW = 1, 2, 4, 8, 16, 32
QK_D9_IBM_ZERO_TARGET = 64
QK_ALL = 127
# BC_flag  (reference src/core/lbm_constants.cuh:377-397)
FLUID, BOUNCE_BACK, ZOU_HE_TOP, ZOU_HE_LEFT = 0, 1, 2, 3
CYLINDER, ZG_OUTFLOW, PRESSURE_OUTLET, REGULARIZED_INLET_TOP = 6, 7, 8, 9
REGULARIZED_BOUNCE_BACK, REGULARIZED_BOUNCE_BACK_CORNER = 11, 12


def build(force=False):
    so = os.path.join(_HERE, "liblbm_oracle.so")
    src = os.path.join(_HERE, "lbm_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liblbm_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        fp = C.POINTER(C.c_float)
        ip = C.POINTER(C.c_int)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, fp, C.c_float,
                                    C.c_float, C.c_float, C.c_int]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_set_flags.argtypes = [C.c_void_p, ip]
        L.oracle_set_markers.argtypes = [C.c_void_p, fp, C.c_int]
        L.oracle_set_marker_velocities.argtypes = [C.c_void_p, fp]
        L.oracle_init.argtypes = [C.c_void_p, fp, fp]
        L.oracle_set_populations.argtypes = [C.c_void_p, fp, fp]
        L.oracle_step.argtypes = [C.c_void_p, C.c_int]
        L.oracle_get_macroscopics.argtypes = [C.c_void_p, fp, fp]
        L.oracle_get_populations.argtypes = [C.c_void_p, fp]
        L.oracle_get_populations_back.argtypes = [C.c_void_p, fp]
        L.oracle_get_force.argtypes = [C.c_void_p, fp]
        L.oracle_get_moment_avg.argtypes = [C.c_void_p, fp]
        L.oracle_total_mass.argtypes = [C.c_void_p]
        L.oracle_total_mass.restype = C.c_double
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_num_threads.argtypes = [C.c_int]
        L.oracle_init_taylor_green.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, fp, fp]
        L.oracle_taylor_green_analytic.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, fp]
        L.oracle_poiseuille_force.argtypes = [C.c_float, C.c_float, C.c_int]
        L.oracle_poiseuille_force.restype = C.c_float
        L.oracle_create_cylinder.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, fp]
        _LIB = L
    return _LIB


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def default_S(coll, omega):
    """S in the row order each operator indexes it (scenario.cuh:47-57 / lidDrivenCavityScenario.cuh:49-59)."""
    if coll >= CM:
        return np.array([0, 0, 0, 1, omega, omega, 1, 1, 1], np.float32)
    return np.array([0, omega, omega, 0, omega, 0, omega, omega, omega], np.float32)


class Oracle:
    """One reference solver instance (LBM<2> + IBMManager<2>) on the CPU."""

    def __init__(self, nx, ny, coll=BGK, viscosity=1.0 / 6.0, S=None, periodic=(True, True), u_max=0.1,
                 force=(0.0, 0.0), quirks=QK_ALL):
        self.nx, self.ny, self.coll = nx, ny, coll
        nu = np.float32(viscosity)
        tau = np.float32(3) * nu + np.float32(0.5)
        self.omega = np.float32(1.0) / tau
        if S is None:
            S = default_S(coll, self.omega)
        self.S = np.ascontiguousarray(S, np.float32)
        self._h = lib().oracle_create(nx, ny, int(periodic[0]), int(periodic[1]), coll, float(nu), _fp(self.S),
                                      float(np.float32(u_max)), float(np.float32(force[0])),
                                      float(np.float32(force[1])), quirks)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_destroy(self._h)
            self._h = None

    def set_flags(self, flags):
        f = np.ascontiguousarray(flags, np.int32).reshape(-1)
        assert f.size == self.nx * self.ny
        lib().oracle_set_flags(self._h, f.ctypes.data_as(C.POINTER(C.c_int)))

    def set_markers(self, pts):
        p = np.ascontiguousarray(pts, np.float32).reshape(-1)
        lib().oracle_set_markers(self._h, _fp(p), p.size // 2)

    def set_marker_velocities(self, vel):
        """IBMBody::velocities [np,2]; honoured only without QK_D9_IBM_ZERO_TARGET (quirks bit 64)."""
        if vel is None:
            lib().oracle_set_marker_velocities(self._h, None)
            return
        v = np.ascontiguousarray(vel, np.float32).reshape(-1)
        lib().oracle_set_marker_velocities(self._h, _fp(v))

    def init(self, rho, u):
        rho = np.ascontiguousarray(rho, np.float32).reshape(-1)
        u = np.ascontiguousarray(u, np.float32).reshape(-1)
        assert rho.size == self.nx * self.ny and u.size == 2 * rho.size
        lib().oracle_init(self._h, _fp(rho), _fp(u))

    def set_populations(self, f, f_back=None):
        f = np.ascontiguousarray(f, np.float32).reshape(-1)
        fb = f if f_back is None else np.ascontiguousarray(f_back, np.float32).reshape(-1)
        lib().oracle_set_populations(self._h, _fp(f), _fp(fb))

    def step(self, n=1):
        lib().oracle_step(self._h, n)

    def macroscopics(self):
        rho = np.empty(self.nx * self.ny, np.float32)
        u = np.empty(2 * self.nx * self.ny, np.float32)
        lib().oracle_get_macroscopics(self._h, _fp(rho), _fp(u))
        return rho.reshape(self.ny, self.nx), u.reshape(self.ny, self.nx, 2)

    def populations(self, back=False):
        f = np.empty(9 * self.nx * self.ny, np.float32)
        (lib().oracle_get_populations_back if back else lib().oracle_get_populations)(self._h, _fp(f))
        return f.reshape(self.ny, self.nx, 9)

    def force(self):
        f = np.empty(2 * self.nx * self.ny, np.float32)
        lib().oracle_get_force(self._h, _fp(f))
        return f.reshape(self.ny, self.nx, 2)

    def moment_avg(self):
        a = np.empty(3, np.float32)
        lib().oracle_get_moment_avg(self._h, _fp(a))
        return a

    def total_mass(self):
        return lib().oracle_total_mass(self._h)


def taylor_green_init(nx, ny, nu=1.0 / 6.0, u_max=0.04):
    rho = np.empty(nx * ny, np.float32)
    u = np.empty(2 * nx * ny, np.float32)
    lib().oracle_init_taylor_green(nx, ny, float(np.float32(nu)), float(np.float32(u_max)), _fp(rho), _fp(u))
    return rho.reshape(ny, nx), u.reshape(ny, nx, 2)


def taylor_green_analytic(nx, ny, nu, u_max, t):
    u = np.empty(2 * nx * ny, np.float32)
    lib().oracle_taylor_green_analytic(nx, ny, float(np.float32(nu)), float(np.float32(u_max)), float(t), _fp(u))
    return u.reshape(ny, nx, 2)


def poiseuille_force(vis, u_max, ny):
    return lib().oracle_poiseuille_force(float(np.float32(vis)), float(np.float32(u_max)), ny)


def create_cylinder(cx, cy, r, num_pts=16):
    p = np.empty(2 * num_pts, np.float32)
    lib().oracle_create_cylinder(float(cx), float(cy), float(r), num_pts, _fp(p))
    return p.reshape(num_pts, 2)


def num_threads():
    return lib().oracle_num_threads()


def set_num_threads(n=None):
    """OpenMP threads of the following oracle calls; None = every core this process may run on (torchrun sets OMP_NUM_THREADS=1)."""
    if n is None:
        n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().oracle_set_num_threads(int(n))
    return num_threads()
