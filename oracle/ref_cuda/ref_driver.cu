// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Driver that runs the *unmodified arithmetic* of the reference CUDA solver
// (Carabalone/cuda-lbm) for its D2Q9 path and dumps rho / u / f so that the
// oracle (oracle/lbm_oracle.c) and the B200-native solver can be pinned to it.
//
// It is compiled by oracle/build_ref.sh against a private, patched copy of the
// reference sources (SURVEY.md Appendix B patch set; the copy lives only in
// oracle/_ref/ which is git-ignored) and linked with the reference's own
// translation units.  Nothing in here is reference code: it only *calls* the
// reference's public LBM<2> methods in the order of src/main.cu:96-114.
//
// Compile-time selection (the reference itself is compile-time configured,
// src/defines.hpp:4-61):
//   -DNX= -DNY= -DSCALE=           grid (src/defines.hpp:20-66)
//   -DREF_CASE=  0 Taylor-Green (periodic XY)            taylorGreenScenario.cuh
//                1 Poiseuille, BB walls, periodic X       poiseuilleScenario.cuh (no IBM body, P9)
//                2 lid-driven cavity, regularized BCs     lidDrivenCavityScenario.cuh
//                3 cavity with ZOU_HE_TOP lid + BOUNCE_BACK walls (older variant, lidDrivenCavityFunctors.cuh:41-53 comments)
//                4 flow past cylinder, IBM markers        flowPastCylinderScenario.cuh (P10)
//                5 flow past cylinder, CYLINDER flag nodes + PRESSURE_OUTLET
//                6 the reference's OWN flowPastCylinderScenario.cuh as it is (two 16-marker cylinders, BGK<2> after the one-token build
//                  fix of SURVEY.md A-D5), for tests/test_shim_gpu.py::test_reference_cylinder_scenario_...
//   -DREF_COLL=  0 BGK<2>  1 MRT<2>  2 CM<2,NoAdapter>  3 CM<2,OptimalAdapter>
//   -DREF_UMAX= -DREF_VISC=        scenario constants (optional)
//   -DREF_NP=                      markers per cylinder (case 4)
//
// Usage: ref_xxx <steps> <outdir> <tag> [dump_step ...]
//   prints one line  "REF_MLUPS <value> steps <n> ms_per_step <t>"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <iostream>
#include <fstream>
#include <sstream>
#include <cmath>
#include <filesystem>
#include <array>
#include <algorithm>
#include <cuda_runtime.h>

#define private public      // to read LBM<2>::d_f for population dumps
#include "core/lbm.cuh"
#undef private
#include "functors/includes.cuh"
#include "IBM/IBMBody.cuh"
#include "IBM/IBM_generators.cuh"
#include "scenarios/scenario.cuh"
#include "scenarios/taylorGreen/taylorGreenFunctors.cuh"
#include "scenarios/poiseuille/poiseuilleFunctors.cuh"
#include "scenarios/lidDrivenCavity/lidDrivenCavityFunctors.cuh"
#include "scenarios/flowPastCylinder/flowPastCylinderFunctors.cuh"

#ifndef REF_CASE
#define REF_CASE 0
#endif
#ifndef REF_COLL
#define REF_COLL 0
#endif

#if REF_COLL == 0
using Coll = BGK<2>;
#elif REF_COLL == 1
using Coll = MRT<2>;
#elif REF_COLL == 2
using Coll = CM<2, NoAdapter>;
#else
using Coll = CM<2, OptimalAdapter>;
#endif

// S in the order each operator indexes it (scenario.cuh:47-57 for MRT rows,
// lidDrivenCavityScenario.cuh:49-59 for CM rows).
#if REF_COLL >= 2
#define REF_S(om) {0.0f, 0.0f, 0.0f, 1.0f, om, om, 1.0f, 1.0f, 1.0f}
#else
#define REF_S(om) {0.0f, om, om, 0.0f, om, 0.0f, om, om, om}
#endif

#if REF_CASE == 6
#include "scenarios/flowPastCylinder/flowPastCylinderScenario.cuh"
using Scenario = FlowPastCylinderScenario;
#elif REF_CASE == 0
#ifndef REF_UMAX
#define REF_UMAX 0.04f
#endif
#ifndef REF_VISC
#define REF_VISC (1.0f/6.0f)
#endif
struct Scenario : public ScenarioTrait<TaylorGreenInit, TaylorGreenBoundary, TaylorGreenValidation, Coll> {
    static constexpr float u_max = REF_UMAX;
    static constexpr float viscosity = REF_VISC;
    static constexpr float tau = viscosity_to_tau(viscosity);
    static constexpr float omega = 1.0f / tau;
#if REF_COLL == 1 && defined(REF_TG_MRT_S)
    // the MRT rates written in taylorGreenScenario.cuh:32-42
    static constexpr float S[quadratures] = {0.0f, 1.0f, 1.4f, 0.0f, 1.2f, 0.0f, 1.9f, omega, omega};
#else
    static constexpr float S[quadratures] = REF_S(omega);
#endif
    static const char* name() { return "TaylorGreen"; }
    static InitType init() { return InitType(viscosity, u_max); }
    static BoundaryType boundary() { return BoundaryType(); }
    static ValidationType validation() { return ValidationType(u_max, viscosity, t); }
};
#elif REF_CASE == 1
#ifndef REF_UMAX
#define REF_UMAX 0.05f
#endif
#ifndef REF_VISC
#define REF_VISC (1.0f/6.0f)
#endif
struct Scenario : public ScenarioTrait<PoiseuilleInit, PoiseuilleBoundary, PoiseuilleValidation, Coll> {
    static constexpr float u_max = REF_UMAX;
    static constexpr float viscosity = REF_VISC;
    static constexpr float tau = viscosity_to_tau(viscosity);
    static constexpr float omega = 1.0f / tau;
    static constexpr float S[quadratures] = REF_S(omega);
    static const char* name() { return "Poiseuille"; }
    static InitType init() { return InitType(u_max); }
    static BoundaryType boundary() { return BoundaryType(); }
    static ValidationType validation() { return ValidationType(u_max, viscosity); }
#ifdef REF_POIS_BODY
    // the immersed cylinder the reference's own Poiseuille scenario adds (poiseuilleScenario.cuh:46-53)
    static void add_bodies() { IBM_bodies.push_back(create_cylinder(48.0f, NY / 2.0f, 8.0f)); }
#endif
};
#elif REF_CASE == 2 || REF_CASE == 3
#ifndef REF_UMAX
#define REF_UMAX 0.1f
#endif
#ifndef REF_VISC
#define REF_VISC 0.0128f
#endif
struct ZouHeLidBoundary {   // the variant left in comments at lidDrivenCavityFunctors.cuh:41-53
    __host__ __device__ int operator()(int x, int y) {
        if ((x == 0 && y == 0) || (x == 0 && y == NY-1) || (x == NX-1 && y == 0) || (x == NX-1 && y == NY-1))
            return BC_flag::BOUNCE_BACK;
        else if (y == NY-1) return BC_flag::ZOU_HE_TOP;
        else if (x == 0 || x == NX-1 || y == 0) return BC_flag::BOUNCE_BACK;
        return BC_flag::FLUID;
    }
};
#if REF_CASE == 2
using LidB = LidDrivenBoundary;
#else
using LidB = ZouHeLidBoundary;
#endif
struct Scenario : public ScenarioTrait<LidDrivenInit, LidB, void, Coll> {
    static constexpr float u_max = REF_UMAX;
    static constexpr float viscosity = REF_VISC;
    static constexpr float tau = viscosity_to_tau(viscosity);
    static constexpr float omega = 1.0f / tau;
    static constexpr float S[quadratures] = REF_S(omega);
    static const char* name() { return "LidDriven"; }
    static InitType init() { return InitType(u_max); }
    static BoundaryType boundary() { return BoundaryType(); }
};
#else
#ifndef REF_UMAX
#define REF_UMAX 0.05f
#endif
#ifndef REF_NP
#define REF_NP 16
#endif
struct CylFlagBoundary {    // CYLINDER flag nodes (boundaries.cuh:58-60) + PRESSURE_OUTLET (:66-68)
    float cx, cy, r;
    CylFlagBoundary(float cx, float cy, float r) : cx(cx), cy(cy), r(r) {}
    __host__ __device__ int operator()(int x, int y) {
        if (y == 0 || y == NY-1) return BC_flag::BOUNCE_BACK;
        if (x == 0) return BC_flag::ZOU_HE_LEFT;
        if (x == NX-1) return BC_flag::PRESSURE_OUTLET;
        float dx = x - cx, dy = y - cy;
        if ((dx * dx + dy * dy) <= (r * r)) return BC_flag::CYLINDER;
        return BC_flag::FLUID;
    }
};
#if REF_CASE == 4
using CylB = FlowPastCylinderBoundary;
#else
using CylB = CylFlagBoundary;
#endif
struct Scenario : public ScenarioTrait<DefaultInit<2>, CylB, void, Coll> {
    static constexpr float Re = 50.0f;
    static constexpr float D = NY / 8.0f;          // 16 at the reference's native NY=128 (flowPastCylinderScenario.cuh:17)
    static constexpr float r = D / 2.0f;
    static constexpr float cx = 3.0f * D;          // 48 at NY=128 (:20)
    static constexpr float cy = NY / 2.0f;
    static constexpr float u_max = REF_UMAX;
#ifdef REF_VISC
    static constexpr float viscosity = REF_VISC;
#else
    static constexpr float viscosity = u_max * D / Re;
#endif
    static constexpr float tau = viscosity_to_tau(viscosity);
    static constexpr float omega = 1.0f / tau;
    static constexpr float S[quadratures] = REF_S(omega);
    static const char* name() { return "FlowPastCylinder"; }
    static InitType init() { return InitType(); }
    static BoundaryType boundary() { return BoundaryType(cx, cy, r); }
#if REF_CASE == 4
    static void add_bodies() { IBM_bodies.push_back(create_cylinder(cx, cy, r, REF_NP)); }
#endif
};
#endif

static void dump(const std::string& path, const void* p, size_t bytes) {
    FILE* fp = fopen(path.c_str(), "wb");
    if (!fp) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(2); }
    fwrite(p, 1, bytes, fp);
    fclose(fp);
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s steps outdir tag [dump_step...]\n", argv[0]); return 1; }
    const int steps = atoi(argv[1]);
    const std::string outdir = argv[2], tag = argv[3];
    std::vector<int> dumps;
    for (int i = 4; i < argc; i++) dumps.push_back(atoi(argv[i]));
    checkCudaErrors(cudaSetDevice(0));

    const size_t N = (size_t)NX * NY;
    LBM<dimensions> lbm;
    lbm.allocate<Scenario>();
    // the reference never initialises d_u outside what the Init functor writes (A-D13) nor d_force before reset
    lbm.init<Scenario>();

    std::vector<float> h_f(N * quadratures);
    auto dump_state = [&](int step) {
        lbm.update_macroscopics();
        checkCudaErrors(cudaMemcpy(h_f.data(), lbm.d_f, N * quadratures * sizeof(float), cudaMemcpyDeviceToHost));
        std::ostringstream base;
        base << outdir << "/" << tag << "_t" << step;
        dump(base.str() + ".rho.bin", lbm.h_rho.data(), N * sizeof(float));
        dump(base.str() + ".u.bin", lbm.h_u.data(), 2 * N * sizeof(float));
        dump(base.str() + ".f.bin", h_f.data(), N * quadratures * sizeof(float));
    };
    if (std::find(dumps.begin(), dumps.end(), 0) != dumps.end()) dump_state(0);   // state right after init<Scenario>()
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float total_ms = 0.0f;
    int timed = 0;
    for (int t = 0; t < steps; t++) {
        cudaEventRecord(e0);
        // order of src/main.cu:96-114
        lbm.increase_ts<Scenario>();
        lbm.stream();
        lbm.swap_buffers();
        lbm.apply_boundaries<Scenario>();
        lbm.uncorrected_macroscopics();
        lbm.reset_forces<Scenario>();
        lbm.ibm_step();
        lbm.correct_macroscopics();
        lbm.compute_equilibrium();
        lbm.collide<Scenario::CollisionOp>();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (t >= 3) { total_ms += ms; timed++; }
        if (std::find(dumps.begin(), dumps.end(), t + 1) != dumps.end()) dump_state(t + 1);
    }
    if (timed > 0) {
        double ms_step = total_ms / timed;
        printf("REF_MLUPS %.3f steps %d ms_per_step %.6f nx %d ny %d case %d coll %d\n",
               (double)N / (ms_step * 1e-3) / 1e6, timed, ms_step, NX, NY, REF_CASE, REF_COLL);
    }
    fflush(stdout);
    return 0;
}
