// functors/includes.cuh — the reference pulls every boundary-condition functor and DefaultInit in through this header
// (src/functors/includes.cuh:4-17).  On this path the boundary functors are not callable objects of the user's
// translation unit: BounceBack, ZouHe, CylinderBoundary, ZG_OutflowBoundary, PressureOutlet, RegularizedInlet,
// RegularizedBounceBack and RegularizedCornerBounceBack are compiled into the fused step kernel
// (cuda_lbm_b200/csrc/d2q9.cuh, apply_bc) and selected per node by the BC_flag the scenario's Boundary functor returns.
#ifndef FUNCTOR_INCLUDES_H
#define FUNCTOR_INCLUDES_H
#include "functors/initialConditions/defaultInit.cuh"
#endif  // FUNCTOR_INCLUDES_H
