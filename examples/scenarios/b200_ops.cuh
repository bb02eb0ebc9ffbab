#pragma once
#include "core/collision/collision.cuh"
// collision operator from an integer (nvcc splits -D values at commas, so a type with two template arguments cannot be
// passed on the command line): 0 BGK<2>, 1 MRT<2>, 2 CM<2,NoAdapter>, 3 CM<2,OptimalAdapter>
template <int id> struct b200_op_by_id;
template <> struct b200_op_by_id<0> { using type = BGK<2>; };
template <> struct b200_op_by_id<1> { using type = MRT<2>; };
template <> struct b200_op_by_id<2> { using type = CM<2, NoAdapter>; };
template <> struct b200_op_by_id<3> { using type = CM<2, OptimalAdapter>; };
