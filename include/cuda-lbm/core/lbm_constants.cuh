// core/lbm_constants.cuh — what scenario code expects from the reference header of the same name
// (src/core/lbm_constants.cuh:1-399), for the D2Q9 path of the B200 engine.
//
// Layout note: the reference's 2-D scenarios index `u[2*node+c]` and
// `u[get_vec_index(node,c)]` interchangeably, which only agree for AoS vectors.  The functor-facing
// arrays of this shim are AoS, so get_vec_index / get_node_index are the AoS forms
// (src/core/lbm_constants.cuh:326-335); the engine's own population storage (SoA, in place) is
// never visible to scenario code.
#ifndef LBM_CONSTANTS_H
#define LBM_CONSTANTS_H

#include <cmath>
#include <cuda_runtime.h>
#include "defines.hpp"

constexpr int dimensions = 2;
constexpr int quadratures = 9;
constexpr float cs = 0.57735026918962f;

// lattice tables in the reference's numbering (src/core/lbm_constants.cuh:13-31): rest, +x, +y, -x, -y, then the diagonals
constexpr float h_weights[quadratures] = {4.0f / 9.0f, 1.0f / 9.0f, 1.0f / 9.0f, 1.0f / 9.0f, 1.0f / 9.0f,
                                          1.0f / 36.0f, 1.0f / 36.0f, 1.0f / 36.0f, 1.0f / 36.0f};
constexpr int h_C[quadratures * dimensions] = {0, 0, 1, 0, 0, 1, -1, 0, 0, -1, 1, 1, -1, 1, -1, -1, 1, -1};
constexpr int h_OPP[quadratures] = {0, 3, 4, 1, 2, 7, 8, 5, 6};

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

// Scenario::viscosity on the device.  PoiseuilleInit::apply_forces reads it (poiseuilleFunctors.cuh:37); LBM<2>::init
// uploads it before any functor runs (the reference does so in send_consts, src/core/lbm.cuh:63-66).  One copy per
// translation unit: instantiate LBM<2>::init<Scenario>() in the unit that defines the scenario (as src/main.cu does).
static __constant__ float vis;

constexpr int num_nodes = NX * NY * NZ;
constexpr int alloc_nodes = num_nodes;

__device__ __host__ __forceinline__ int get_node_index(int node, int quadrature = 0) { return node * quadratures + quadrature; }
__device__ __host__ __forceinline__ int get_vec_index(int node, int component) { return node * dimensions + component; }
__device__ __host__ __forceinline__ int get_node_from_coords(int x, int y, int z = 0) { (void)z; return y * NX + x; }
__device__ __host__ __forceinline__ void get_coords_from_node(int node, int& x, int& y, int& z) {
    x = node % NX;
    y = node / NX;
    z = 0;
}

constexpr inline float viscosity_to_tau(float v) { return 3 * v + 0.5f; }
constexpr inline float tau_to_viscosity(float t) { return (t - 0.5f) / 3.0f; }
constexpr float compute_reynolds(float u_max, float domain_size, float viscosity) { return (u_max * domain_size) / viscosity; }

// Node classes a Boundary functor may return.  The numeric values are the interface (they travel through
// lbm_set_flags as LBM_* codes, include/lbm_b200.h), so the enumerator order of src/core/lbm_constants.cuh:377-397 is kept.
// The engine implements the 2-D members; a 3-D-only class is rejected by lbm_set_flags' caller below.
enum BC_flag {
    FLUID, BOUNCE_BACK, ZOU_HE_TOP, ZOU_HE_LEFT, ZOU_HE_TOP_LEFT_TOP_INFLOW, ZOU_HE_TOP_RIGHT_TOP_INFLOW, CYLINDER, ZG_OUTFLOW,
    PRESSURE_OUTLET, REGULARIZED_INLET_TOP, REGULARIZED_INLET_LEFT, REGULARIZED_BOUNCE_BACK, REGULARIZED_BOUNCE_BACK_CORNER,
    EXTRAPOLATED_CORNER_EDGE, CORNER_EDGE_BOUNCE_BACK, PRESSURE_INLET_LEFT, GUO_VELOCITY_INLET, GUO_PRESSURE_OUTLET, REGULARIZED_OUTLET
};

#endif  // LBM_CONSTANTS_H
