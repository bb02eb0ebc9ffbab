// Taylor-Green vortex on a doubly periodic box — BASELINE.json configs 1 and 4.
// Same physics and metric as the reference's scenario (src/scenarios/taylorGreen/taylorGreenFunctors.cuh:25-47,66-81,
// taylorGreenScenario.cuh:59-88): vortex evaluated at cell centres, rho = 1 + 3 P, u0 = 0.04 / SCALE, nu = 1/6, L2 error
// of u against the analytically decaying field in percent.  Written against include/cuda-lbm/scenarios/scenario.cuh.
//   -DB200_TG_OP=1 (MRT<2>), 2 (CM<2,NoAdapter>), 3 (CM<2,OptimalAdapter>) selects the collision operator; default 0, BGK<2>.
#pragma once
#include <cmath>
#include <vector>
#include "scenarios/scenario.cuh"
#include "scenarios/b200_ops.cuh"

#ifndef B200_TG_OP
#define B200_TG_OP 0      // 0 BGK<2>, 1 MRT<2>, 2 CM<2,NoAdapter>, 3 CM<2,OptimalAdapter>
#endif

namespace b200_tg {
struct Field {
    float u0, nu;
    __host__ __device__ void at(int xi, int yi, float time, float& rho, float& ux, float& uy) const {
        const float x = xi + 0.5f, y = yi + 0.5f;
        const float kx = 2.0f * (float)M_PI / NX, ky = 2.0f * (float)M_PI / NY;
        const float decay = expf(-time * nu * (kx * kx + ky * ky));
        ux = -u0 * sqrtf(ky / kx) * cosf(kx * x) * sinf(ky * y) * decay;
        uy = u0 * sqrtf(kx / ky) * sinf(kx * x) * cosf(ky * y) * decay;
        const float P = -0.25f * u0 * u0 * ((ky / kx) * cosf(2 * kx * x) + (kx / ky) * cosf(2 * ky * y)) * decay * decay;
        rho = 1.0f + 3.0f * P;
    }
};
}  // namespace b200_tg

struct B200TaylorGreenInit {
    b200_tg::Field f;
    B200TaylorGreenInit(float nu, float u_max) : f{u_max / SCALE, nu} {}
    __host__ __device__ void apply_forces(float* rho, float* u, float* force, int node) {
        force[get_vec_index(node, 0)] = 0.0f;
        force[get_vec_index(node, 1)] = 0.0f;
    }
    __host__ __device__ void operator()(float* rho, float* u, float* force, int node) {
        float r, ux, uy;
        f.at(node % NX, node / NX, 0.0f, r, ux, uy);
        rho[node] = r;
        u[get_vec_index(node, 0)] = ux;
        u[get_vec_index(node, 1)] = uy;
        apply_forces(rho, u, force, node);
    }
};

struct B200AllFluid {
    __host__ __device__ int operator()(int, int) const { return BC_flag::FLUID; }
};

struct B200TaylorGreenValidation {
    b200_tg::Field f;
    float t;
    __host__ __device__ B200TaylorGreenValidation(float u_max, float nu, float t) : f{u_max / SCALE, nu}, t(t) {}
    __host__ __device__ void operator()(int x, int y, float& ux, float& uy) const { float r; f.at(x, y, t, r, ux, uy); }
};

struct B200TaylorGreenScenario : public ScenarioTrait<B200TaylorGreenInit, B200AllFluid, B200TaylorGreenValidation, b200_op_by_id<B200_TG_OP>::type> {
    static constexpr float u_max = 0.04f;
    static constexpr float viscosity = 1.0f / 6.0f;
    static constexpr float tau = viscosity_to_tau(viscosity);
    static constexpr float omega = 1.0f / tau;
    static constexpr bool periodic_x = true, periodic_y = true;
    // rates in the row order of the selected operator: MRT rows (rho,e,eps,jx,qx,jy,qy,pxx,pxy) or CM rows (rho,kx,ky,bulk,shear,shear,h.o. x3)
    static constexpr float S[quadratures] = {0.0f,
                                             CollisionOp::lbm_b200_op >= LBM_CM ? 0.0f : omega,
                                             CollisionOp::lbm_b200_op >= LBM_CM ? 0.0f : omega,
                                             CollisionOp::lbm_b200_op >= LBM_CM ? 1.0f : 0.0f,
                                             omega,
                                             CollisionOp::lbm_b200_op >= LBM_CM ? omega : 0.0f,
                                             CollisionOp::lbm_b200_op >= LBM_CM ? 1.0f : omega,
                                             CollisionOp::lbm_b200_op >= LBM_CM ? 1.0f : omega,
                                             CollisionOp::lbm_b200_op >= LBM_CM ? 1.0f : omega};
    static const char* name() { return "TaylorGreen"; }
    static InitType init() { return InitType(viscosity, u_max); }
    static BoundaryType boundary() { return BoundaryType(); }
    static ValidationType validation() { return ValidationType(u_max, viscosity, t); }

    // 100 * || u - u_analytic ||_2 / || u_analytic ||_2 over the whole grid
    template <typename LBMSolver>
    static float compute_error(LBMSolver& solver) {
        if (solver.update_ts < solver.timestep) solver.update_macroscopics();
        const auto exact = validation();
        double err = 0.0, ref = 0.0;
        for (int y = 0; y < NY; y++)
            for (int x = 0; x < NX; x++) {
                float ax, ay;
                exact(x, y, ax, ay);
                const size_t i = ((size_t)y * NX + x) * dimensions;
                const double dx = solver.h_u[i] - ax, dy = solver.h_u[i + 1] - ay;
                err += dx * dx + dy * dy;
                ref += (double)ax * ax + (double)ay * ay;
            }
        return (float)(std::sqrt(err / ref) * 100.0);
    }
};
