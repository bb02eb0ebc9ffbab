// Lid-driven cavity, Re = u_max NY / nu = 1000, CM<2,OptimalAdapter>, regularized walls and lid — BASELINE.json config 3.
// Boundary classes as in the reference's cavity (src/scenarios/lidDrivenCavity/lidDrivenCavityFunctors.cuh:38-57):
// corners REGULARIZED_BOUNCE_BACK_CORNER, lid (y = NY-1) REGULARIZED_INLET_TOP moving at u_max in +x, remaining walls
// REGULARIZED_BOUNCE_BACK; CM relaxation set of lidDrivenCavityScenario.cuh:49-59.  No closed-form solution: the
// reference validates against Ghia et al.'s centre-line tables at 129^2; this file is the throughput configuration.
#pragma once
#include "scenarios/scenario.cuh"
#include "scenarios/b200_ops.cuh"

#ifndef B200_LID_OP
#define B200_LID_OP 3      // 0 BGK<2>, 1 MRT<2>, 2 CM<2,NoAdapter>, 3 CM<2,OptimalAdapter>
#endif
#ifndef B200_LID_RE
#define B200_LID_RE 1000.0f
#endif

struct B200CavityWalls {
    __host__ __device__ int operator()(int x, int y) const {
        const bool ex = (x == 0 || x == NX - 1), ey = (y == 0 || y == NY - 1);
        if (ex && ey) return BC_flag::REGULARIZED_BOUNCE_BACK_CORNER;
        if (y == NY - 1) return BC_flag::REGULARIZED_INLET_TOP;
        if (ex || y == 0) return BC_flag::REGULARIZED_BOUNCE_BACK;
        return BC_flag::FLUID;
    }
};

struct B200LidDrivenScenario : public ScenarioTrait<DefaultInit<2>, B200CavityWalls, void, b200_op_by_id<B200_LID_OP>::type> {
    static constexpr float u_max = 0.1f;
    static constexpr float viscosity = u_max * NY / B200_LID_RE;
    static constexpr float tau = viscosity_to_tau(viscosity);
    static constexpr float omega = 1.0f / tau;
    static constexpr float S[quadratures] = {0.0f, 0.0f, 0.0f, 1.0f, omega, omega, 1.0f, 1.0f, 1.0f};
    static const char* name() { return "LidDriven"; }
    static InitType init() { return InitType(); }
    static BoundaryType boundary() { return BoundaryType(); }
};
