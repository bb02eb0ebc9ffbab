"""`Engine` — one handle of the C ABI (include/lbm_b200.h) with numpy in / out: what the parity tests, bench.py and the
multi-process launcher (slab.py) drive.  The reference-facing host interface is C++ (include/cuda-lbm/: LBM<2>, ScenarioTrait);
a Python mirror of that driver protocol, used by a development tool only, lives in tools/pyscenarios.py.
"""
import ctypes as C

import numpy as np

from . import _capi as capi
from ._capi import LbmConfig, LbmInfo, check, lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def default_S(collision, omega):
    """Scenario::S in the row order the operator indexes it (reference scenario.cuh:47-57 for BGK/MRT,
    lidDrivenCavityScenario.cuh:49-59 for CM)."""
    om = np.float32(omega)
    if collision >= capi.CM:
        return np.array([0, 0, 0, 1, om, om, 1, 1, 1], np.float32)
    return np.array([0, om, om, 0, om, 0, om, om, om], np.float32)


class Engine:
    def __init__(self, nx, ny, collision=capi.BGK, viscosity=1.0 / 6.0, S=None, periodic=(True, True), u_max=0.1,
                 force=(0.0, 0.0), quirks=capi.QK_REFERENCE, adapter_mode=capi.ADAPTER_EXACT, device=0, rank=0, world=1,
                 ibm_mailbox_nodes=0):
        L = lib()
        cfg = LbmConfig()
        check(L.lbm_default_config(C.byref(cfg)))
        cfg.nx, cfg.ny = nx, ny
        cfg.periodic_x, cfg.periodic_y = int(periodic[0]), int(periodic[1])
        cfg.collision = collision
        cfg.viscosity = float(np.float32(viscosity))
        tau = np.float32(3) * np.float32(viscosity) + np.float32(0.5)
        self.omega = np.float32(1.0) / tau
        if S is None:
            S = default_S(collision, self.omega)
        for i, s in enumerate(np.asarray(S, np.float32)):
            cfg.S[i] = float(s)
        cfg.u_max = float(np.float32(u_max))
        cfg.force_x, cfg.force_y = float(np.float32(force[0])), float(np.float32(force[1]))
        cfg.quirks, cfg.adapter_mode = quirks, adapter_mode
        cfg.device, cfg.rank, cfg.world = device, rank, world
        cfg.ibm_mailbox_nodes = ibm_mailbox_nodes
        self.cfg = cfg
        self._h = C.c_void_p()
        check(L.lbm_create(C.byref(cfg), C.byref(self._h)))
        self.nx, self.ny = nx, ny
        inf = self.info()
        self.y0, self.ny_local = inf.y0, inf.ny_local

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().lbm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        o = LbmInfo()
        check(lib().lbm_info(self._h, C.byref(o)))
        return o

    def set_stream(self, cuda_stream_ptr):
        check(lib().lbm_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def set_flags(self, flags):
        f = np.ascontiguousarray(flags, np.int32).reshape(-1)
        if f.size != self.nx * self.ny:
            raise ValueError("flags must cover the global grid")
        check(lib().lbm_set_flags(self._h, f.ctypes.data_as(C.POINTER(C.c_int32))))

    def set_body_force(self, fx, fy):
        check(lib().lbm_set_body_force(self._h, float(np.float32(fx)), float(np.float32(fy))))

    def set_force_field(self, force):
        if force is None:
            check(lib().lbm_set_force_field(self._h, None))
            return
        f = np.ascontiguousarray(force, np.float32).reshape(-1)
        check(lib().lbm_set_force_field(self._h, _fp(f)))

    def add_body(self, points):
        p = np.ascontiguousarray(points, np.float32).reshape(-1)
        check(lib().lbm_add_body(self._h, _fp(p), p.size // 2))

    def set_body_velocities(self, body, velocities):
        if velocities is None:
            check(lib().lbm_set_body_velocities(self._h, body, None))
            return
        v = np.ascontiguousarray(velocities, np.float32).reshape(-1)
        check(lib().lbm_set_body_velocities(self._h, body, _fp(v)))

    def move_body(self, body, points):
        p = np.ascontiguousarray(points, np.float32).reshape(-1)
        check(lib().lbm_move_body(self._h, body, _fp(p)))

    def init_fields(self, rho, u):
        rho = np.ascontiguousarray(rho, np.float32).reshape(-1)
        u = np.ascontiguousarray(u, np.float32).reshape(-1)
        if rho.size != self.nx * self.ny or u.size != 2 * rho.size:
            raise ValueError("rho/u must cover the global grid")
        check(lib().lbm_init_fields(self._h, _fp(rho), _fp(u)))

    def init_fields_local(self, rho_ptr, u_ptr):
        """rho,u of this slab's rows from (ideally pinned) host memory given as integer addresses; asynchronous."""
        check(lib().lbm_init_fields_local(self._h, C.c_void_p(rho_ptr), C.c_void_p(u_ptr)))

    def reserve_macroscopics(self):
        check(lib().lbm_reserve_macroscopics(self._h))

    def init_taylor_green(self, nu, u0):
        check(lib().lbm_init_taylor_green(self._h, float(np.float32(nu)), float(np.float32(u0))))

    def set_populations(self, f, f_back=None):
        f = np.ascontiguousarray(f, np.float32).reshape(-1)
        fb = None if f_back is None else np.ascontiguousarray(f_back, np.float32).reshape(-1)
        check(lib().lbm_set_populations(self._h, _fp(f), None if fb is None else _fp(fb)))

    def populations(self):
        """Post-collision populations of this slab's rows, [ny_local, nx, 9]."""
        f = np.zeros(self.nx * self.ny * 9, np.float32)
        check(lib().lbm_get_populations(self._h, _fp(f)))
        return f.reshape(self.ny, self.nx, 9)[self.y0:self.y0 + self.ny_local]

    def step(self, n=1, macroscopics=False):
        fn = lib().lbm_step_with_macroscopics if macroscopics else lib().lbm_step
        check(fn(self._h, n))

    def sync(self):
        check(lib().lbm_sync(self._h))

    def macroscopics(self):
        rho = np.empty(self.nx * self.ny_local, np.float32)
        u = np.empty(2 * self.nx * self.ny_local, np.float32)
        check(lib().lbm_get_macroscopics(self._h, rho.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p)))
        return rho.reshape(self.ny_local, self.nx), u.reshape(self.ny_local, self.nx, 2)

    def run_from_host(self, rho_ptr, u_ptr, nsteps, rho_out_ptr, u_out_ptr):
        """One driver segment: host rho,u (this slab's rows) -> nsteps -> host rho,u; integer addresses of (ideally pinned) host memory."""
        check(lib().lbm_run_from_host(self._h, C.c_void_p(rho_ptr), C.c_void_p(u_ptr), nsteps, C.c_void_p(rho_out_ptr), C.c_void_p(u_out_ptr)))

    def macroscopics_into(self, rho_ptr, u_ptr):
        """D2H into caller-owned (ideally pinned) host memory given as integer addresses."""
        check(lib().lbm_get_macroscopics(self._h, C.c_void_p(rho_ptr), C.c_void_p(u_ptr)))

    def total_mass(self):
        m = C.c_double()
        check(lib().lbm_total_mass(self._h, C.byref(m)))
        return m.value

    def moment_avg(self):
        a = np.empty(3, np.float32)
        check(lib().lbm_moment_avg(self._h, _fp(a)))
        return a

    def moment_sums(self):
        s = (C.c_double * 3)()
        check(lib().lbm_get_moment_sums(self._h, s))
        return np.array(list(s))

    def set_moment_sums(self, sums):
        s = (C.c_double * 3)(*[float(v) for v in sums])
        check(lib().lbm_set_moment_sums(self._h, s))

    def adapter_prepass(self):
        check(lib().lbm_adapter_prepass(self._h))

    def adapter_sums_pending(self):
        return bool(lib().lbm_adapter_sums_pending(self._h))

    def recover_macroscopics(self):
        check(lib().lbm_recover_macroscopics(self._h))

    # --- validation on the device (this slab's sums; add them over slabs) ---
    def velocity_error_sums(self, d_u_ref_ptr):
        o = (C.c_double * 2)()
        check(lib().lbm_velocity_error_sums(self._h, C.c_void_p(d_u_ref_ptr), o))
        return np.array([o[0], o[1]])

    def taylor_green_error_sums(self, nu, u0, t):
        o = (C.c_double * 2)()
        check(lib().lbm_taylor_green_error_sums(self._h, float(np.float32(nu)), float(np.float32(u0)), float(np.float32(t)), o))
        return np.array([o[0], o[1]])

    def row_mean_velocity(self):
        mx, my = np.empty(self.ny_local, np.float64), np.empty(self.ny_local, np.float64)
        dp = C.POINTER(C.c_double)
        check(lib().lbm_row_mean_velocity(self._h, mx.ctypes.data_as(dp), my.ctypes.data_as(dp)))
        return mx, my

    def sample_velocity(self, nodes, out=None):
        """u at the listed global nodes, [n, 2]; entries of nodes another slab owns keep the values of `out` (zeros by default)."""
        nodes = np.ascontiguousarray(nodes, np.int64).reshape(-1)
        out = np.zeros((nodes.size, 2), np.float32) if out is None else np.ascontiguousarray(out, np.float32)
        check(lib().lbm_sample_velocity(self._h, nodes.ctypes.data_as(C.POINTER(C.c_int64)), nodes.size, _fp(out)))
        return out

    # --- checkpoint / restart ---
    def checkpoint_bytes(self):
        n = C.c_int64()
        check(lib().lbm_checkpoint_bytes(self._h, C.byref(n)))
        return n.value

    def checkpoint_write(self, path):
        check(lib().lbm_checkpoint_write(self._h, str(path).encode()))

    def checkpoint_read(self, path):
        check(lib().lbm_checkpoint_read(self._h, str(path).encode()))

    # --- peer-mapped neighbours ---
    def peer_export(self):
        buf = (C.c_ubyte * capi.PEER_DESC_BYTES)()
        check(lib().lbm_peer_export(self._h, buf))
        return bytes(buf)

    def peer_attach(self, side, desc):
        buf = (C.c_ubyte * capi.PEER_DESC_BYTES).from_buffer_copy(desc)
        check(lib().lbm_peer_attach(self._h, side, buf))

    def peer_attach_all(self, descs):
        """descs: the peer_export() bytes of ALL slabs, indexed by rank."""
        buf = (C.c_ubyte * (capi.PEER_DESC_BYTES * len(descs))).from_buffer_copy(b"".join(descs))
        check(lib().lbm_peer_attach_all(self._h, buf, len(descs)))

    def peer_detach(self):
        check(lib().lbm_peer_detach(self._h))

    # --- bodies across slab faces, halo coupling (device pointers as ints) ---
    def ibm_exchange_floats(self):
        n = C.c_int64()
        check(lib().lbm_ibm_exchange_floats(self._h, C.byref(n)))
        return n.value

    def ibm_pack(self, ptr):
        check(lib().lbm_ibm_pack(self._h, C.c_void_p(ptr)))

    def ibm_unpack(self, ptr):
        check(lib().lbm_ibm_unpack(self._h, C.c_void_p(ptr)))

    # --- slab halos (device pointers as ints) ---
    def next_step_needs_halo(self):
        return bool(lib().lbm_next_step_needs_halo(self._h))

    def halo(self, what, side, ptr):
        fn = getattr(lib(), "lbm_halo_" + what)
        check(fn(self._h, side, C.c_void_p(ptr)))
