// Channel flow past an immersed cylinder, Re = u_max D / nu = 50, MRT<2>, direct-forcing IBM markers — BASELINE.json config 5.
// Geometry scaled from the reference's scenario (src/scenarios/flowPastCylinder/flowPastCylinderScenario.cuh:15-26,
// flowPastCylinderFunctors.cuh:54-77): D = NY/8, centre (3 D, NY/2); y-walls BOUNCE_BACK, inlet ZOU_HE_LEFT at u_max,
// outlet ZG_OUTFLOW; markers from create_cylinder with -DB200_CYL_NP (default: unit arc-length spacing).
#pragma once
#include "scenarios/scenario.cuh"
#include "scenarios/b200_ops.cuh"
#include "IBM/IBM_generators.cuh"

#ifndef B200_CYL_OP
#define B200_CYL_OP 1      // 0 BGK<2>, 1 MRT<2>, 2 CM<2,NoAdapter>, 3 CM<2,OptimalAdapter>
#endif

struct B200ChannelInOut {
    __host__ __device__ int operator()(int x, int y) const {
        if (y == 0 || y == NY - 1) return BC_flag::BOUNCE_BACK;
        if (x == 0) return BC_flag::ZOU_HE_LEFT;
        if (x == NX - 1) return BC_flag::ZG_OUTFLOW;
        return BC_flag::FLUID;
    }
};

struct B200CylinderScenario : public ScenarioTrait<DefaultInit<2>, B200ChannelInOut, void, b200_op_by_id<B200_CYL_OP>::type> {
    static constexpr float Re = 50.0f;
    static constexpr float D = NY / 8.0f, r = D / 2.0f;
    static constexpr float cx = 3.0f * D, cy = NY / 2.0f;
    static constexpr float u_max = 0.05f;
    static constexpr float viscosity = u_max * D / Re;
    static constexpr float tau = viscosity_to_tau(viscosity);
    static constexpr float omega = 1.0f / tau;
    static constexpr float S[quadratures] = DEFAULT_MRT_S_MATRIX(omega);
#ifdef B200_CYL_NP
    static constexpr int num_markers = B200_CYL_NP;
#else
    static constexpr int num_markers = (int)(2.0f * 3.14159265f * r + 0.5f);
#endif
    static const char* name() { return "FlowPastCylinder"; }
    static InitType init() { return InitType(); }
    static BoundaryType boundary() { return BoundaryType(); }
    static void add_bodies() { IBM_bodies.push_back(create_cylinder(cx, cy, r, num_markers)); }
};
