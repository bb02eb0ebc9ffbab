// Body-force driven channel between two wet-node bounce-back walls, x-periodic — BASELINE.json config 2.
// Physics of the reference's scenario (src/scenarios/poiseuille/poiseuilleFunctors.cuh:20-75, poiseuilleScenario.cuh:12-77)
// with the operator the config names (MRT<2>) and without the immersed cylinder the reference's add_bodies() inserts
// (SURVEY.md Appendix A-D15).  F_x = 8 nu u_max / NY^2.  Two error metrics are reported: the reference's
// (parabola y (NY - y), walls assumed at y = 0 and y = NY) and the profile for walls ON the first / last node row,
// y (NY-1 - y), which is what wet-node bounce-back produces.
#pragma once
#include <cmath>
#include <vector>
#include "scenarios/scenario.cuh"
#include "scenarios/b200_ops.cuh"

#ifndef B200_POIS_OP
#define B200_POIS_OP 1      // 0 BGK<2>, 1 MRT<2>, 2 CM<2,NoAdapter>, 3 CM<2,OptimalAdapter>
#endif

struct B200PoiseuilleInit {
    float u_max;
    B200PoiseuilleInit(float u_max) : u_max(u_max) {}
    __host__ __device__ void apply_forces(float* rho, float* u, float* force, int node) {
#ifdef __CUDA_ARCH__
        const float nu = vis;               // Scenario::viscosity as uploaded by LBM<2>::init
#else
        const float nu = 0.0f;
#endif
        force[get_vec_index(node, 0)] = 8.0f * nu * u_max / (NY * NY);
        force[get_vec_index(node, 1)] = 0.0f;
    }
    __host__ __device__ void operator()(float* rho, float* u, float* force, int node) {
        rho[node] = 1.0f;
        u[get_vec_index(node, 0)] = 0.0f;
        u[get_vec_index(node, 1)] = 0.0f;
        apply_forces(rho, u, force, node);
    }
};

struct B200ChannelWalls {
    __host__ __device__ int operator()(int x, int y) const { return (y == 0 || y == NY - 1) ? BC_flag::BOUNCE_BACK : BC_flag::FLUID; }
};

struct B200PoiseuilleValidation {
    float u_max, nu;
    B200PoiseuilleValidation(float u_max, float nu) : u_max(u_max), nu(nu) {}
    float reference_profile(int y) const { return ((8.0f * nu * u_max / (NY * NY)) / (2.0f * nu)) * y * (NY - y); }
    float wet_node_profile(int y) const { return ((8.0f * nu * u_max / (NY * NY)) / (2.0f * nu)) * y * (NY - 1 - y); }
};

struct B200PoiseuilleScenario : public ScenarioTrait<B200PoiseuilleInit, B200ChannelWalls, B200PoiseuilleValidation, b200_op_by_id<B200_POIS_OP>::type> {
    static constexpr float u_max = 0.05f;
    static constexpr float viscosity = 1.0f / 6.0f;
    static constexpr float tau = viscosity_to_tau(viscosity);
    static constexpr float omega = 1.0f / tau;
    static constexpr bool periodic_x = true, periodic_y = false;
    static constexpr float S[quadratures] = DEFAULT_MRT_S_MATRIX(omega);
    static const char* name() { return "Poiseuille"; }
    static InitType init() { return InitType(u_max); }
    static BoundaryType boundary() { return BoundaryType(); }
    static ValidationType validation() { return ValidationType(u_max, viscosity); }

    // x-averaged u_x against a profile, RMS over rows, in percent of u_max (the reference's metric when wet == false)
    template <typename LBMSolver>
    static float profile_error(LBMSolver& solver, bool wet) {
        if (solver.update_ts < solver.timestep) solver.update_macroscopics();
        const auto v = validation();
        double sum = 0.0;
        for (int y = 0; y < NY; y++) {
            double ux = 0.0;
            for (int x = 0; x < NX; x++) ux += solver.h_u[((size_t)y * NX + x) * dimensions];
            const double d = ux / NX - (wet ? v.wet_node_profile(y) : v.reference_profile(y));
            sum += d * d;
        }
        return (float)(std::sqrt(sum / NY) * 100.0 / u_max);
    }
    template <typename LBMSolver>
    static float compute_error(LBMSolver& solver) { return profile_error(solver, false); }
};
