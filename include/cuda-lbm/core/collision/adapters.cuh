// core/collision/adapters.cuh — adapter tags of CM<2,Adapter> (reference src/core/collision/adapters.cuh:7-22,48-111).
// The arithmetic of OptimalAdapter lives in the engine (cuda_lbm_b200/csrc/d2q9.cuh, optimal_rate); here the types only
// select it.
#ifndef ADAPTERS_H
#define ADAPTERS_H

struct AdapterBase {
    static constexpr int i_star = 5;     // moments with index > i_star take the adapter's rate (adapters.cuh:8-13)
};
struct NoAdapter : AdapterBase { static constexpr bool lbm_b200_optimal = false; };
struct OptimalAdapter : AdapterBase { static constexpr bool lbm_b200_optimal = true; };

#endif  // ADAPTERS_H
