// core/collision/collision.cuh — the collision-operator template arguments of ScenarioTrait and LBM<2>::collide<Op>()
// (reference src/core/collision/collision.cuh:16-65, BGK/BGK.cuh, MRT/MRT.cuh, CM/CM.cuh).
// In the reference these types carry the per-node `apply`; here they are tags naming the fused kernel the engine runs
// (one register-resident stream+collide kernel per operator, cuda_lbm_b200/csrc/kernels.cuh).
#ifndef COLLISION_OPS_H
#define COLLISION_OPS_H

#include "core/lbm_constants.cuh"
#include "core/collision/adapters.cuh"
#include "../../../lbm_b200.h"

template <int dim>
struct BGK {
    static_assert(dim == 2, "D2Q9 path only");
    static constexpr int lbm_b200_op = LBM_BGK;
};
template <int dim>
struct MRT {
    static_assert(dim == 2, "D2Q9 path only");
    static constexpr int lbm_b200_op = LBM_MRT;
};
template <int dim, typename AdapterType = NoAdapter>
struct CM {
    static_assert(dim == 2, "D2Q9 path only");
    static constexpr int lbm_b200_op = AdapterType::lbm_b200_optimal ? LBM_CM_OPTIMAL : LBM_CM;
};

#endif  // COLLISION_OPS_H
