"""DEVELOPMENT TOOLING (used by tools/config_bench.py only) — a Python/numpy mirror of the reference's scenario layer
(src/scenarios/, L4 in SURVEY.md §1) and of the LBM<2> driver protocol over cuda_lbm_b200.Engine.  The reference-facing host interface
of this repository is the C++ header shim include/cuda-lbm/ (LBM<2>, ScenarioTrait); bench.py measures the BASELINE configurations
through it (examples/main.cu).

`ScenarioTrait` keeps the members of src/scenarios/scenario.cuh:22-78 (InitType / BoundaryType /
ValidationType / CollisionOp as class attributes, viscosity / tau / omega / S / u_max, name(), init(),
boundary(), validation(), add_bodies(), compute_error(solver), update_ts(t), IBM_bodies,
has_analytical_solution).  Grid size, which the reference bakes in as NX/NY macros
(src/defines.hpp:20-66), is passed to the factory methods instead.  Functors are vectorised numpy
versions of the reference's per-node functors, in float32.
"""
import math

import numpy as np

from cuda_lbm_b200 import _capi as capi
from cuda_lbm_b200.solver import Engine

f32 = np.float32


def viscosity_to_tau(v):            # lbm_constants.cuh:365-367
    return f32(3) * f32(v) + f32(0.5)


def tau_to_viscosity(t):            # lbm_constants.cuh:369-371
    return (f32(t) - f32(0.5)) / f32(3.0)


def compute_reynolds(u_max, domain_size, viscosity):    # lbm_constants.cuh:373-375
    return f32(u_max) * f32(domain_size) / f32(viscosity)


def DEFAULT_MRT_S_MATRIX(omega):    # scenario.cuh:10-20
    om = f32(omega)
    return np.array([0, om, om, 0, om, 0, om, om, om], f32)


def create_cylinder(cx, cy, r, num_pts=16):
    """IBMBody::points of create_cylinder — src/IBM/IBM_generators.cu:5-25 (AoS [i*2+c], float32)."""
    angle = f32(2 * math.pi / num_pts)
    i = np.arange(num_pts, dtype=f32)
    pts = np.empty((num_pts, 2), f32)
    pts[:, 0] = f32(cx) + f32(r) * np.cos(i * angle, dtype=f32)
    pts[:, 1] = f32(cy) + f32(r) * np.sin(i * angle, dtype=f32)
    return pts


class DefaultInit:
    """DefaultInit<2> — src/functors/initialConditions/defaultInit.cuh:4-31."""

    def __init__(self, nx, ny):
        self.nx, self.ny = nx, ny

    def apply_forces(self):
        return (0.0, 0.0)

    def __call__(self):
        return np.ones((self.ny, self.nx), f32), np.zeros((self.ny, self.nx, 2), f32)


class FluidBoundary:
    def __call__(self, x, y):
        return np.zeros_like(x, dtype=np.int32)


class ScenarioTrait:
    InitType = DefaultInit
    BoundaryType = FluidBoundary
    ValidationType = None
    collision = capi.BGK                         # CollisionOp = BGK<2>  (scenario.cuh:26)
    viscosity = f32(1.0 / 6.0)
    u_max = f32(0.1)
    periodic = (False, False)                    # PERIODIC_X / PERIODIC_Y of streaming.cuh:8-11
    t = 0.0
    S = None                                     # None -> the default of scenario.cuh:47-57 for the operator

    def __init__(self):
        self.IBM_bodies = []
        self.tau = viscosity_to_tau(self.viscosity)
        self.omega = f32(1.0) / self.tau
        if self.S is None:
            from cuda_lbm_b200.solver import default_S
            self.S = default_S(self.collision, self.omega)
        self.S = np.asarray(self.S, f32)

    has_analytical_solution = property(lambda self: self.ValidationType is not None)

    def name(self):
        return "BaseScenario"

    def init(self, nx, ny):
        return self.InitType(nx, ny)

    def boundary(self, nx, ny):
        return self.BoundaryType()

    def body_force(self, nx, ny):
        """What Init::apply_forces writes into d_force for every node every step (macroscopics.cuh:13-48)."""
        return self.init(nx, ny).apply_forces()

    def add_bodies(self, nx, ny):
        return

    def update_ts(self, new_ts):
        self.t = float(new_ts)

    def compute_error(self, solver):
        raise NotImplementedError


# ---------------------------------------------------------------- Taylor-Green (src/scenarios/taylorGreen/)
class TaylorGreenInit:
    """taylorGreenFunctors.cuh:7-48"""

    def __init__(self, nx, ny, nu, u_max, scale):
        self.nx, self.ny, self.nu = nx, ny, f32(nu)
        self.u_max = f32(u_max) / f32(scale)

    def apply_forces(self):
        return (0.0, 0.0)

    def __call__(self):
        nx, ny = self.nx, self.ny
        x = (np.arange(nx, dtype=f32) + f32(0.5))[None, :]
        y = (np.arange(ny, dtype=f32) + f32(0.5))[:, None]
        kx, ky = f32(2.0 * math.pi / nx), f32(2.0 * math.pi / ny)
        um = self.u_max
        ux = -um * np.sqrt(ky / kx, dtype=f32) * np.cos(kx * x, dtype=f32) * np.sin(ky * y, dtype=f32)
        uy = um * np.sqrt(kx / ky, dtype=f32) * np.sin(kx * x, dtype=f32) * np.cos(ky * y, dtype=f32)
        P = f32(-0.25) * um * um * ((ky / kx) * np.cos(f32(2) * kx * x, dtype=f32) + (kx / ky) * np.cos(f32(2) * ky * y, dtype=f32))
        rho = (f32(1.0) + f32(3.0) * P).astype(f32) * np.ones((ny, nx), f32)
        u = np.stack([ux * np.ones((ny, nx), f32), uy * np.ones((ny, nx), f32)], axis=-1).astype(f32)
        return rho.astype(f32), u


class TaylorGreenValidation:
    """taylorGreenFunctors.cuh:57-93"""

    def __init__(self, nx, ny, u0, nu, t, scale):
        self.nx, self.ny, self.nu, self.t, self.scale = nx, ny, f32(nu), f32(t), scale

    def getFullField(self):
        nx, ny = self.nx, self.ny
        x = (np.arange(nx, dtype=f32) + f32(0.5))[None, :]
        y = (np.arange(ny, dtype=f32) + f32(0.5))[:, None]
        u_max = f32(0.04) / f32(self.scale)                # hard-coded at :72
        kx, ky = f32(2.0 * math.pi / nx), f32(2.0 * math.pi / ny)
        td = f32(1.0) / (self.nu * (kx * kx + ky * ky))
        decay = np.exp(-self.t / td, dtype=f32)
        ux = -u_max * np.sqrt(ky / kx, dtype=f32) * np.cos(kx * x, dtype=f32) * np.sin(ky * y, dtype=f32) * decay
        uy = u_max * np.sqrt(kx / ky, dtype=f32) * np.sin(kx * x, dtype=f32) * np.cos(ky * y, dtype=f32) * decay
        return np.stack([ux * np.ones((ny, nx), f32), uy * np.ones((ny, nx), f32)], axis=-1).astype(f32)


class TaylorGreenScenario(ScenarioTrait):
    """taylorGreenScenario.cuh:7-88 (BGK<2>, u_max 0.04, nu 1/6, periodic X and Y, SCALE = NX/128)"""
    ValidationType = TaylorGreenValidation
    collision = capi.BGK
    u_max = f32(0.04)
    viscosity = f32(1.0 / 6.0)
    periodic = (True, True)

    def __init__(self, scale=1, collision=None):
        if collision is not None:
            self.collision = collision
        super().__init__()
        self.scale = scale

    def name(self):
        return "TaylorGreen"

    def init(self, nx, ny):
        return TaylorGreenInit(nx, ny, self.viscosity, self.u_max, self.scale)

    def validation(self, nx, ny):
        return TaylorGreenValidation(nx, ny, self.u_max, self.viscosity, self.t, self.scale)

    def compute_error(self, solver):                       # taylorGreenScenario.cuh:59-88
        ana = self.validation(solver.NX, solver.NY).getFullField()
        if solver.update_ts < solver.timestep:
            solver.update_macroscopics()
        u = solver.h_u.reshape(solver.NY, solver.NX, 2)
        err = np.sum((u - ana).astype(np.float64) ** 2)
        norm = np.sum(ana.astype(np.float64) ** 2)
        return float(np.sqrt(err / norm) * 100.0)


# ---------------------------------------------------------------- Poiseuille (src/scenarios/poiseuille/)
class PoiseuilleInit:
    """poiseuilleFunctors.cuh:7-49"""

    def __init__(self, nx, ny, u_max, vis):
        self.nx, self.ny, self.u_max, self.vis = nx, ny, f32(u_max), f32(vis)

    def apply_forces(self):
        return (float(f32(8.0) * self.vis * self.u_max / f32(self.ny * self.ny)), 0.0)     # :37

    def __call__(self):
        return np.ones((self.ny, self.nx), f32), np.zeros((self.ny, self.nx, 2), f32)


class PoiseuilleBoundary:
    """poiseuilleFunctors.cuh:52-61"""

    def __init__(self, ny):
        self.ny = ny

    def __call__(self, x, y):
        return np.where((y == 0) | (y == self.ny - 1), capi.BOUNCE_BACK, capi.FLUID).astype(np.int32)


class PoiseuilleValidation:
    """poiseuilleFunctors.cuh:63-86"""

    def __init__(self, ny, u_max, viscosity):
        self.ny, self.u_max, self.viscosity = ny, f32(u_max), f32(viscosity)

    def getProfile(self):
        y = np.arange(self.ny, dtype=f32)
        ny = f32(self.ny)
        return ((f32(8.0) * self.viscosity * self.u_max / f32(self.ny * self.ny)) / (f32(2.0) * self.viscosity)) * y * (ny - y)


class PoiseuilleScenario(ScenarioTrait):
    """poiseuilleScenario.cuh:8-77.  The reference file adds an IBM cylinder to the channel
    (:46-53, SURVEY A-D15); `with_body=False` (default) is BASELINE config 2."""
    ValidationType = PoiseuilleValidation
    collision = capi.BGK
    u_max = f32(0.05)
    viscosity = f32(1.0 / 6.0)
    periodic = (True, False)

    def __init__(self, collision=None, with_body=False):
        if collision is not None:
            self.collision = collision
        super().__init__()
        self.with_body = with_body

    def name(self):
        return "Poiseuille"

    def init(self, nx, ny):
        return PoiseuilleInit(nx, ny, self.u_max, self.viscosity)

    def boundary(self, nx, ny):
        return PoiseuilleBoundary(ny)

    def validation(self, nx, ny):
        return PoiseuilleValidation(ny, self.u_max, self.viscosity)

    def add_bodies(self, nx, ny):
        if self.with_body:
            self.IBM_bodies.append(create_cylinder(48.0, ny / 2.0, 8.0))

    def compute_error(self, solver):                       # poiseuilleScenario.cuh:55-77
        prof = self.validation(solver.NX, solver.NY).getProfile()
        if solver.update_ts < solver.timestep:
            solver.update_macroscopics()
        ux = solver.h_u.reshape(solver.NY, solver.NX, 2)[:, :, 0]
        avg = ux.sum(axis=1, dtype=np.float32) / f32(solver.NX)
        return float(np.sqrt(np.sum((avg - prof).astype(np.float64) ** 2) / solver.NY) * 100.0 / float(self.u_max))


# ---------------------------------------------------------------- lid-driven cavity (src/scenarios/lidDrivenCavity/)
class LidDrivenBoundary:
    """lidDrivenCavityFunctors.cuh:36-58"""

    def __init__(self, nx, ny):
        self.nx, self.ny = nx, ny

    def __call__(self, x, y):
        nx, ny = self.nx, self.ny
        corner = ((x == 0) | (x == nx - 1)) & ((y == 0) | (y == ny - 1))
        f = np.zeros_like(x, dtype=np.int32)
        f[(x == 0) | (x == nx - 1) | (y == 0)] = capi.REGULARIZED_BOUNCE_BACK
        f[y == ny - 1] = capi.REGULARIZED_INLET_TOP
        f[corner] = capi.REGULARIZED_BOUNCE_BACK_CORNER
        return f


class LidDrivenValidation:
    """Ghia, Ghia & Shin (1982) centre-line tables as used at lidDrivenCavityFunctors.cuh:60-221
    (public data of the paper; 17 stations, Re = 100, 400, 1000 kept here)."""
    ghia_y = np.array([0.0, 0.0547, 0.0625, 0.0703, 0.1016, 0.1719, 0.2813, 0.4531, 0.5, 0.6172, 0.7344, 0.8516,
                       0.9531, 0.9609, 0.9688, 0.9766, 1.0], f32)
    ghia_x = np.array([0.0, 0.0625, 0.0703, 0.0781, 0.0938, 0.1563, 0.2266, 0.2344, 0.5, 0.8047, 0.8594, 0.9063,
                       0.9453, 0.9531, 0.9609, 0.9688, 1.0], f32)
    ux = {
        100: [0.0, -0.03717, -0.04192, -0.04775, -0.06434, -0.1015, -0.15662, -0.2109, -0.20581, -0.13641, 0.00332,
              0.23151, 0.68717, 0.73722, 0.78871, 0.84123, 1.0],
        400: [0.0, -0.08186, -0.09266, -0.10338, -0.14612, -0.24299, -0.32726, -0.17119, -0.11477, 0.02135, 0.16256,
              0.29093, 0.55892, 0.61756, 0.68439, 0.75837, 1.0],
        1000: [0.0, -0.18109, -0.20196, -0.2222, -0.2973, -0.38289, -0.27805, -0.10648, -0.0608, 0.05702, 0.18719,
               0.33304, 0.46604, 0.51117, 0.57492, 0.65928, 1.0],
    }
    uy = {
        100: [0.0, 0.09233, 0.10091, 0.1089, 0.12317, 0.16077, 0.17507, 0.17527, 0.05454, -0.24533, -0.22445,
              -0.16914, -0.10313, -0.08864, -0.07391, -0.05906, 0.0],
        400: [0.0, 0.1836, 0.19713, 0.2092, 0.22965, 0.28124, 0.30203, 0.30174, 0.05188, -0.38598, -0.44993,
              -0.23827, -0.22847, -0.19254, -0.15663, -0.12146, 0.0],
        1000: [0.0, 0.27485, 0.29012, 0.30353, 0.32627, 0.37095, 0.33075, 0.32235, 0.02526, -0.31966, -0.42665,
               -0.5155, -0.39188, -0.33714, -0.27669, -0.21388, 0.0],
    }

    def get_closest_ref_data(self, re, is_ux):             # :201-221
        avail = sorted(self.ux.keys())
        closest = min(avail, key=lambda a: abs(re - a))
        return np.array((self.ux if is_ux else self.uy)[closest], f32)


class LidDrivenScenario(ScenarioTrait):
    """lidDrivenCavityScenario.cuh:9-157 (regularized boundaries, CM rates)"""
    ValidationType = LidDrivenValidation
    collision = capi.BGK
    u_max = f32(0.0517)
    viscosity = f32(0.0667)
    periodic = (False, False)

    def __init__(self, collision=None, u_max=None, viscosity=None):
        if collision is not None:
            self.collision = collision
        if u_max is not None:
            self.u_max = f32(u_max)
        if viscosity is not None:
            self.viscosity = f32(viscosity)
        tau = viscosity_to_tau(self.viscosity)
        om = f32(1.0) / tau
        self.S = np.array([0, 0, 0, 1, om, om, 1, 1, 1], f32)      # the CM-ordered S of :49-59, whatever the operator
        super().__init__()

    def name(self):
        return "LidDriven"

    def boundary(self, nx, ny):
        return LidDrivenBoundary(nx, ny)

    def validation(self, nx=None, ny=None):
        return LidDrivenValidation()

    def compute_error(self, solver):                       # lidDrivenCavityScenario.cuh:88-157
        if solver.update_ts < solver.timestep:
            solver.update_macroscopics()
        v = self.validation()
        NX, NY = solver.NX, solver.NY
        re = int(compute_reynolds(self.u_max, NY, self.viscosity))
        ux_ref, uy_ref = v.get_closest_ref_data(re, True), v.get_closest_ref_data(re, False)
        u = solver.h_u.reshape(NY, NX, 2)
        cxn, cyn = NX // 2, NY // 2
        eu = ru = ev = rv = 0.0
        for i in range(17):
            y = int(min(max(round(float(v.ghia_y[i]) * (NY - 1)), 0), NY - 1))
            d = float(u[y, cxn, 0]) / float(self.u_max) - float(ux_ref[i])
            eu += d * d
            ru += float(ux_ref[i]) ** 2
            x = int(min(max(round(float(v.ghia_x[i]) * (NX - 1)), 0), NX - 1))
            d = float(u[cyn, x, 1]) / float(self.u_max) - float(uy_ref[i])
            ev += d * d
            rv += float(uy_ref[i]) ** 2
        r1 = math.sqrt(eu / ru) if ru > 0 else 0.0
        r2 = math.sqrt(ev / rv) if rv > 0 else 0.0
        return 100.0 * (r1 + r2) / 2.0


# ---------------------------------------------------------------- flow past cylinder (src/scenarios/flowPastCylinder/)
class FlowPastCylinderBoundary:
    """flowPastCylinderFunctors.cuh:48-77"""

    def __init__(self, nx, ny):
        self.nx, self.ny = nx, ny

    def __call__(self, x, y):
        f = np.zeros_like(x, dtype=np.int32)
        f[x == self.nx - 1] = capi.ZG_OUTFLOW
        f[x == 0] = capi.ZOU_HE_LEFT
        f[(y == 0) | (y == self.ny - 1)] = capi.BOUNCE_BACK
        return f


class FlowPastCylinderScenario(ScenarioTrait):
    """flowPastCylinderScenario.cuh:8-77 with the geometry scaled by NY (D = NY/8, cx = 3D; 16 and 48 at the
    reference's native NY = 128), u = 0 everywhere initially (A-D13), one IBM cylinder of `num_pts` markers."""
    collision = capi.BGK
    u_max = f32(0.05)
    Re = f32(50.0)
    periodic = (False, False)

    def __init__(self, ny, collision=None, num_pts=16):
        if collision is not None:
            self.collision = collision
        self.D = f32(ny / 8.0)
        self.r = self.D / f32(2.0)
        self.cx, self.cy = f32(3.0) * self.D, f32(ny / 2.0)
        self.viscosity = self.u_max * self.D / self.Re
        self.num_pts = num_pts
        super().__init__()

    def name(self):
        return "FlowPastCylinder"

    def boundary(self, nx, ny):
        return FlowPastCylinderBoundary(nx, ny)

    def add_bodies(self, nx, ny):
        self.IBM_bodies.append(create_cylinder(self.cx, self.cy, self.r, self.num_pts))


class LBM:
    """Drop-in for the reference's `LBM<2>` object (see module docstring).  `Scenario` objects follow
    cuda_lbm_b200.scenarios.ScenarioTrait, the mirror of src/scenarios/scenario.cuh:22-78."""

    def __init__(self, nx, ny, device=0, quirks=capi.QK_REFERENCE, adapter_mode=capi.ADAPTER_EXACT):
        self.NX, self.NY = nx, ny
        self.device, self.quirks, self.adapter_mode = device, quirks, adapter_mode
        self.timestep = 0
        self.update_ts = 0
        self.h_rho = np.zeros(nx * ny, np.float32)
        self.h_u = np.zeros(2 * nx * ny, np.float32)
        self.engine = None
        self._pending = False

    # LBM::allocate<Scenario>()  lbm.cuh:92-125
    def allocate(self, S):
        self.engine = Engine(self.NX, self.NY, collision=S.collision, viscosity=S.viscosity, S=S.S, periodic=S.periodic,
                             u_max=S.u_max, force=S.body_force(self.NX, self.NY), quirks=self.quirks,
                             adapter_mode=self.adapter_mode, device=self.device)
        S.IBM_bodies.clear()
        S.add_bodies(self.NX, self.NY)
        for body in S.IBM_bodies:
            self.engine.add_body(body)

    # LBM::init<Scenario>()  init.cuh:45-86
    def init(self, S):
        init = S.init(self.NX, self.NY)
        rho, u = init()
        boundary = S.boundary(self.NX, self.NY)
        yy, xx = np.meshgrid(np.arange(self.NY), np.arange(self.NX), indexing="ij")
        self.engine.set_flags(boundary(xx, yy))
        self.engine.init_fields(rho, u)
        self.timestep = 0
        self._pending = False

    def _close_step(self, macroscopics=False):
        if self._pending:
            self.engine.step(1, macroscopics=macroscopics)
            self._pending = False

    def increase_ts(self, S=None):
        self._close_step()
        self.timestep += 1
        if S is not None:
            S.update_ts(self.timestep)

    # the reference's per-kernel host methods (lbm.cuh:345-377): recorded, executed fused
    def stream(self): pass
    def swap_buffers(self): pass
    def apply_boundaries(self, S=None): pass
    def uncorrected_macroscopics(self): pass
    def reset_forces(self, S=None): pass
    def ibm_step(self): pass
    def correct_macroscopics(self): pass
    def compute_equilibrium(self): pass

    def collide(self, op=None):
        self._pending = True

    def run(self, nsteps, S=None):
        """nsteps iterations of the main.cu loop body without the per-call overhead."""
        self._close_step()
        if nsteps > 0:
            self.engine.step(nsteps - 1)
            self.timestep += nsteps
            self._pending = True
            if S is not None:
                S.update_ts(self.timestep)

    # LBM::update_macroscopics()  lbm.cuh:148-154
    def update_macroscopics(self):
        if self._pending:
            self._close_step(macroscopics=True)
        rho, u = self.engine.macroscopics()
        self.h_rho[:] = rho.reshape(-1)
        self.h_u[:] = u.reshape(-1)
        self.update_ts = self.timestep

    def get_rho(self):
        return self.h_rho

    def get_u(self):
        return self.h_u

    # LBM::compute_error<Scenario>()  lbm.cuh:163-171
    def compute_error(self, S):
        if S.has_analytical_solution:
            return S.compute_error(self)
        print("Scenario does not provide verification/validation.")
        return 0.0

    def free(self):
        if self.engine is not None:
            self.engine.close()
            self.engine = None
