// Lid-driven cavity validated against Ghia, Ghia & Shin (1982) at a selectable Reynolds number and collision operator.
// The centre-line tables are the reference's (LidDrivenValidation, src/scenarios/lidDrivenCavity/lidDrivenCavityFunctors.cuh:58-221):
// this file INCLUDES the reference's functor header and is therefore only built where the reference tree is present
// (examples/Makefile, target `ref`); its Init / Boundary functors are the reference's too.  What it adds is the choice the
// reference makes by editing its scenario file (lidDrivenCavityScenario.cuh:16-42): -DB200_GHIA_RE=100|400|1000|..., -DB200_GHIA_OP=0..3.
#pragma once
#include "scenarios/scenario.cuh"
#include "scenarios/lidDrivenCavity/lidDrivenCavityFunctors.cuh"
#include "scenarios/b200_ops.cuh"

#ifndef B200_GHIA_RE
#define B200_GHIA_RE 1000
#endif
#ifndef B200_GHIA_OP
#define B200_GHIA_OP 2      // CM<2,NoAdapter>
#endif

struct B200GhiaCavityScenario : public ScenarioTrait<LidDrivenInit, LidDrivenBoundary, LidDrivenValidation, b200_op_by_id<B200_GHIA_OP>::type> {
    static constexpr float u_max = 0.1f;
    static constexpr float viscosity = u_max * NY / (float)B200_GHIA_RE;
    static constexpr float tau = viscosity_to_tau(viscosity);
    static constexpr float omega = 1.0f / tau;
    // CM rows (rho, kx, ky, bulk, shear, shear, h.o. x3), lidDrivenCavityScenario.cuh:49-59; MRT / BGK: the BGK-equivalent set
    static constexpr float S[quadratures] = {0.0f,
                                             CollisionOp::lbm_b200_op >= LBM_CM ? 0.0f : omega,
                                             CollisionOp::lbm_b200_op >= LBM_CM ? 0.0f : omega,
                                             CollisionOp::lbm_b200_op >= LBM_CM ? 1.0f : 0.0f,
                                             omega,
                                             CollisionOp::lbm_b200_op >= LBM_CM ? omega : 0.0f,
                                             CollisionOp::lbm_b200_op >= LBM_CM ? 1.0f : omega,
                                             CollisionOp::lbm_b200_op >= LBM_CM ? 1.0f : omega,
                                             CollisionOp::lbm_b200_op >= LBM_CM ? 1.0f : omega};
    static const char* name() { return "LidDrivenGhia"; }
    static InitType init() { return InitType(u_max); }
    static BoundaryType boundary() { return BoundaryType(); }
    static ValidationType validation() { return ValidationType(); }
    // the reference's metric (lidDrivenCavityScenario.cuh:88-157) with the samples gathered on the device
    template <typename LBMSolver>
    static float compute_error(LBMSolver& solver) { return solver.template centerline_error_device<B200GhiaCavityScenario>(); }
};
