// lanes.cuh — the arithmetic types of the per-cell functions.
//
// V1 holds one cell, V2 two neighbouring cells in a 64-bit register pair.  sm_100a has packed fp32 instructions
// (FADD2 / FMUL2 / FFMA2: two IEEE fp32 results per issue slot); the vectorised step kernel processes its four cells as
// two V2 lanes, which halves the floating-point instruction count of an otherwise issue-limited kernel.
// Every operation is an explicitly rounded intrinsic (no compiler contraction), so V1 and V2 evaluate a formula with
// bit-identical results per cell: a cell gets the same value whether the scalar or the vector kernel computes it.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace lbm {

#define LBM_HD __host__ __device__ __forceinline__
#ifdef __CUDA_ARCH__
#define LBM_UNROLL _Pragma("unroll")
#else
#define LBM_UNROLL
#endif
// Host builds (tests/host_math_check.cu runs the same templates on the CPU) use plain IEEE operations; x86-64 without
// -mfma does not contract, and fmaf() rounds once, so host and device agree bit for bit.
#ifndef __CUDA_ARCH__
LBM_HD float h_add(float a, float b) { return a + b; }
LBM_HD float h_mul(float a, float b) { return a * b; }
#define __fadd_rn(a, b) h_add(a, b)
#define __fsub_rn(a, b) h_add(a, -(b))
#define __fmul_rn(a, b) h_mul(a, b)
#define __fmaf_rn(a, b, c) fmaf(a, b, c)
#define __fadd2_rn(a, b) make_float2(h_add((a).x, (b).x), h_add((a).y, (b).y))
#define __fmul2_rn(a, b) make_float2(h_mul((a).x, (b).x), h_mul((a).y, (b).y))
#define __ffma2_rn(a, b, c) make_float2(fmaf((a).x, (b).x, (c).x), fmaf((a).y, (b).y, (c).y))
#endif

// Approximate square root / reciprocal for the OptimalAdapter's sensor quantities (rho|u|, |Pi|, 1 / (3 tau* + 1/2), 1 / grid mean): one
// MUFU instruction each (<= 1 ulp) instead of the ~10-instruction IEEE sequences.  They only steer the relaxation rate of the three
// highest central moments (tau* = 8.7e-3 +- O(1e-3) x these ratios, adapters.cuh:55-109): a relative 1e-7 there moves a population by
// less than 1e-10.  Everything a conserved or hydrodynamic moment depends on (1 / rho, the transforms) stays IEEE.  Host builds
// (tests/host_math_check.cu) use the exact functions.
LBM_HD float fast_sqrt(float x) {
#ifdef __CUDA_ARCH__
    float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
    return sqrtf(x);
#endif
}
LBM_HD float fast_rcp(float x) {
#ifdef __CUDA_ARCH__
    float r; asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
    return 1.0f / x;
#endif
}

struct V1 { float a; };
struct V2 { float2 a; };

LBM_HD float2 f2(float s) { return make_float2(s, s); }
LBM_HD float2 neg2(float2 v) { return make_float2(-v.x, -v.y); }

// ---- V1
LBM_HD V1 operator+(V1 x, V1 y) { return {__fadd_rn(x.a, y.a)}; }
LBM_HD V1 operator-(V1 x, V1 y) { return {__fsub_rn(x.a, y.a)}; }
LBM_HD V1 operator*(V1 x, V1 y) { return {__fmul_rn(x.a, y.a)}; }
LBM_HD V1 operator*(V1 x, float s) { return {__fmul_rn(x.a, s)}; }
LBM_HD V1 operator*(float s, V1 x) { return {__fmul_rn(x.a, s)}; }
LBM_HD V1 operator+(V1 x, float s) { return {__fadd_rn(x.a, s)}; }
LBM_HD V1 operator-(V1 x) { return {-x.a}; }
LBM_HD V1 fma(V1 x, V1 y, V1 z) { return {__fmaf_rn(x.a, y.a, z.a)}; }
LBM_HD V1 fma(V1 x, float s, V1 z) { return {__fmaf_rn(x.a, s, z.a)}; }
LBM_HD V1 fma(float s, V1 y, V1 z) { return {__fmaf_rn(s, y.a, z.a)}; }
LBM_HD V1 fma(V1 x, float s, float t) { return {__fmaf_rn(x.a, s, t)}; }
LBM_HD V1 fma(V1 x, V1 y, float t) { return {__fmaf_rn(x.a, y.a, t)}; }
LBM_HD V1 rcp(V1 x) { return {1.0f / x.a}; }               // IEEE division, as the reference's 1/rho
LBM_HD V1 vsqrt(V1 x) { return {fast_sqrt(x.a)}; }
LBM_HD void bcast(V1& out, float s) { out.a = s; }
LBM_HD float hsum(V1 x) { return x.a; }

// ---- V2
LBM_HD V2 operator+(V2 x, V2 y) { return {__fadd2_rn(x.a, y.a)}; }
LBM_HD V2 operator-(V2 x, V2 y) { return {__fadd2_rn(x.a, neg2(y.a))}; }     // operand negation is an instruction modifier
LBM_HD V2 operator*(V2 x, V2 y) { return {__fmul2_rn(x.a, y.a)}; }
LBM_HD V2 operator*(V2 x, float s) { return {__fmul2_rn(x.a, f2(s))}; }
LBM_HD V2 operator*(float s, V2 x) { return {__fmul2_rn(x.a, f2(s))}; }
LBM_HD V2 operator+(V2 x, float s) { return {__fadd2_rn(x.a, f2(s))}; }
LBM_HD V2 operator-(V2 x) { return {neg2(x.a)}; }
LBM_HD V2 fma(V2 x, V2 y, V2 z) { return {__ffma2_rn(x.a, y.a, z.a)}; }
LBM_HD V2 fma(V2 x, float s, V2 z) { return {__ffma2_rn(x.a, f2(s), z.a)}; }
LBM_HD V2 fma(float s, V2 y, V2 z) { return {__ffma2_rn(f2(s), y.a, z.a)}; }
LBM_HD V2 fma(V2 x, float s, float t) { return {__ffma2_rn(x.a, f2(s), f2(t))}; }
LBM_HD V2 fma(V2 x, V2 y, float t) { return {__ffma2_rn(x.a, y.a, f2(t))}; }
LBM_HD V2 rcp(V2 x) { return {make_float2(1.0f / x.a.x, 1.0f / x.a.y)}; }
LBM_HD V2 vsqrt(V2 x) { return {make_float2(fast_sqrt(x.a.x), fast_sqrt(x.a.y))}; }
LBM_HD void bcast(V2& out, float s) { out.a = f2(s); }
LBM_HD float hsum(V2 x) { return x.a.x + x.a.y; }

template <class V> LBM_HD V splat(float s) { V v; bcast(v, s); return v; }

}  // namespace lbm
