// scenarios/scenario.cuh — the ScenarioTrait contract (reference src/scenarios/scenario.cuh:10-78).
//
// A scenario derives from ScenarioTrait<Init, Boundary, Validation, CollisionOp, Adapter>, shadows the constants it
// changes (viscosity / tau / omega / S / u_max), and provides init(), boundary(), validation(), and optionally
// add_bodies() and compute_error(solver).  Every member the reference's trait has exists here under the same name, so
// the scenario files of src/scenarios/ compile against this header unchanged.
//
// Additions a scenario MAY make (detected, never required):
//   static constexpr bool periodic_x, periodic_y;   // instead of the PERIODIC_X / PERIODIC_Y macros (core/streaming/streaming.cuh)
#ifndef FLOW_SCENARIO_H
#define FLOW_SCENARIO_H

#include <array>
#include <string>
#include <type_traits>
#include <vector>
#include "core/collision/collision.cuh"
#include "IBM/IBMBody.cuh"
#include "IBM/IBM_generators.cuh"

// relaxation rates in Lallemand-Luo row order (rho, e, eps, jx, qx, jy, qy, pxx, pxy) that make MRT<2> coincide with BGK at rate omega_val
#define DEFAULT_MRT_S_MATRIX(omega_val) {0.0f, omega_val, omega_val, 0.0f, omega_val, 0.0f, omega_val, omega_val, omega_val}

template <typename InitFunctor, typename BoundaryFunctor, typename ValidationFunctor = void, typename CollisionType = BGK<2>,
          typename AdapterType = NoAdapter>
struct ScenarioTrait {
    using InitType = InitFunctor;
    using BoundaryType = BoundaryFunctor;
    using ValidationType = ValidationFunctor;
    using CollisionOp = CollisionType;
    using AdapterOp = AdapterType;

    // defaults: nu = 1/6 -> tau = 1, omega = 1
    static constexpr float viscosity = 1.0f / 6.0f;
    static constexpr float tau = viscosity_to_tau(viscosity);
    static constexpr float omega = 1.0f / tau;
    static constexpr float u_max = 0.1f;
    // A scenario that selects MRT<2> or CM<2,...> should shadow S in the row order of its operator; this default is the
    // BGK-equivalent MRT set.
    static constexpr float S[quadratures] = DEFAULT_MRT_S_MATRIX(omega);

    static inline float t = 0.0f;                       // simulation time, advanced by LBM<2>::increase_ts
    static inline std::vector<IBMBody> IBM_bodies;      // filled by add_bodies(), consumed by LBM<2>::allocate

    static constexpr bool has_analytical_solution = !std::is_same<ValidationFunctor, void>::value;

    static const char* name() { return "BaseScenario"; }
    static InitType initCondition() { return InitType(); }
    static BoundaryType boundaryCondition() { return BoundaryType(); }

    template <typename LBMSolver>
    static float compute_error(LBMSolver& solver);

    static void add_bodies() {}
    static void update_ts(float new_ts) { t = new_ts; }
};

#endif  // FLOW_SCENARIO_H
