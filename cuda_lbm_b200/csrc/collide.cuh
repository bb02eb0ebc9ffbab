// collide.cuh — lattice constants, moments and the four collision operators of the D2Q9 path, written once over the lane
// types of lanes.cuh (V1 = one cell, V2 = two cells in packed fp32 instructions).
//
// Reference formulas (file:line under src/ of Carabalone/cuda-lbm):
//   moments    core/macroscopics/macroscopics.cu:5-38         rho, u* = sum f c / rho, Pi_ab = sum f c_a c_b
//   BGK        core/collision/BGK/BGK.cuh:13-51               f -= omega (f - f_eq) ; + Guo force term
//   f_eq       core/equilibrium/equilibrium.cu:5-39           w rho (1 + 3 c.u + 4.5 (c.u)^2 - 1.5 u^2)
//   MRT        core/collision/MRT/MRT.cu:4-76, M / M^-1 of core/lbm_constants.cuh:33-55
//   CM         core/collision/CM/CM.cuh:27-250                central moments about u, T^-1(u)
//   adapter    core/collision/adapters.cuh:48-111
// The algebra is the reference's; the evaluation order is not (the 9x9 transforms are folded into add / FMA chains with
// their 0, +-1, +-2 entries resolved at compile time, opposite directions share their symmetric part), so results agree
// with the reference to fp32 round-off — tests/host_math_check.cu checks every operator against a direct fp64
// evaluation of the definitions, tests/test_parity_gpu.py against the CPU restatement of the reference.
#pragma once
#include "lanes.cuh"

namespace lbm {

constexpr int Q = 9;

// quirk bits (include/lbm_b200.h)
constexpr int QK_D1 = 1, QK_D2 = 2, QK_D3 = 4, QK_D7 = 8, QK_D8 = 16, QK_D11 = 32, QK_D9 = 64;

// lattice tables — reference src/core/lbm_constants.cuh:13-31 (h_C, h_OPP, h_weights)
__host__ __device__ __forceinline__ constexpr int cx(int q) { return (q == 1 || q == 5 || q == 8) ? 1 : ((q == 3 || q == 6 || q == 7) ? -1 : 0); }
__host__ __device__ __forceinline__ constexpr int cy(int q) { return (q == 2 || q == 5 || q == 6) ? 1 : ((q == 4 || q == 7 || q == 8) ? -1 : 0); }
__host__ __device__ __forceinline__ constexpr int opp(int q) { return q == 0 ? 0 : (q <= 4 ? ((q + 1) % 4) + 1 : ((q - 3) % 4) + 5); }
__host__ __device__ __forceinline__ constexpr float wq(int q) { return q == 0 ? 4.0f / 9.0f : (q <= 4 ? 1.0f / 9.0f : 1.0f / 36.0f); }
static_assert(opp(1) == 3 && opp(2) == 4 && opp(3) == 1 && opp(4) == 2 && opp(5) == 7 && opp(6) == 8 && opp(7) == 5 && opp(8) == 6, "OPP");

// what the collision operators need from the scenario: Scenario::omega, Scenario::S, the quirk mask
struct Relax {
    float omega;
    float S[Q];
    int quirks;
};

template <class V>
struct Mom { V rho, inv_rho, ux, uy, pxx, pxy, pyy; };

template <class V>
LBM_HD Mom<V> moments_v(const V g[Q]) {
    Mom<V> m;
    m.rho = (((((((g[0] + g[1]) + g[2]) + g[3]) + g[4]) + g[5]) + g[6]) + g[7]) + g[8];
    const V d56 = g[5] - g[6], d87 = g[8] - g[7], d58 = g[5] - g[8], d67 = g[6] - g[7];
    const V jx = ((g[1] - g[3]) + d56) + d87;
    const V jy = ((g[2] - g[4]) + d58) + d67;
    m.inv_rho = rcp(m.rho);
    m.ux = jx * m.inv_rho;
    m.uy = jy * m.inv_rho;
    const V d = ((g[5] + g[6]) + g[7]) + g[8];
    m.pxx = (g[1] + g[3]) + d;
    m.pyy = (g[2] + g[4]) + d;
    m.pxy = d56 - d87;                      // (g5 - g6) + (g7 - g8)
    return m;
}

// |Pi| = sqrt(Pxx^2 + 2 Pxy^2 + Pyy^2) — macroscopics.cu:30-32
template <class V>
LBM_HD V pi_norm_v(const Mom<V>& m) { return vsqrt(fma(m.pyy, m.pyy, fma(m.pxy * 2.0f, m.pxy, m.pxx * m.pxx))); }

// rho |u| — the second adapter quantity (macroscopics.cuh:51-120), one definition for every kernel that sums or uses it
template <class V>
LBM_HD V jmag_v(V ux, V uy, V rho) { return vsqrt(fma(ux, ux, uy * uy)) * rho; }

// ------------------------------------------------------------------ BGK with Guo forcing
// Directions come in opposite pairs (q, opp q) with c.u of opposite sign:  f_eq(+-) = w rho (1 + e2) +- 3 w rho cu with the
// second-order part e2 = 4.5 cu^2 - 1.5 u^2, and the Guo term (+-) = w k (9 cu cF - 3 u.F) +- 3 w k cF with k = 1 - omega/2.  The relaxation
// f' = (1 - omega) f + omega f_eq + force is evaluated as one FMA per direction on the shared symmetric / antisymmetric parts.
// e2 (~Ma^2) is kept apart from the 1 and enters through an FMA, omega w rho (1 + e2) = fma(orw, e2, orw): formed as 1 + e2 first, its
// rounding (6e-8, the same for all nine directions of a cell) is a mass error per cell and step that acts as pressure noise — at
// |u| ~ 6e-4 (8192^2 Taylor-Green) it tripled the analytic error relative to the reference, whose f_eq bracket is evaluated in double
// (equilibrium.cu:7-37, A-D19); with the FMA form the engine is as accurate (tests/test_reference_fullsize_gpu.py).
template <class V>
LBM_HD void bgk_pair(V& ga, V& gb, V cu, V cF, V m15usq, V m3uF, V orw, float onem, float k, float w, bool forced) {
    // ga: direction with +cu, gb: its opposite.  orw = omega * w * rho
    const V e2 = fma(cu * 4.5f, cu, m15usq);
    V sym = fma(orw, e2, orw);
    V anti = (orw * 3.0f) * cu;
    if (forced) {
        sym = fma(fma(cu * 9.0f, cF, m3uF), w * k, sym);
        anti = fma(cF, 3.0f * w * k, anti);
    }
    ga = fma(ga, onem, sym + anti);
    gb = fma(gb, onem, sym - anti);
}

template <class V>
LBM_HD void collide_bgk_v(const Relax& r, V g[Q], V rho, V ux, V uy, bool forced, V Fx, V Fy) {
    const float om = r.omega, onem = 1.0f - om, k = 1.0f - 0.5f * om;
    const V m15usq = fma(ux, ux, uy * uy) * -1.5f;
    const V orho = rho * om;
    V m3uF = splat<V>(0.0f);
    if (forced) m3uF = fma(ux, Fx, uy * Fy) * -3.0f;
    // rest population: cu = cF = 0
    {
        const V orw0 = orho * (4.0f / 9.0f);
        V s = fma(orw0, m15usq, orw0);
        if (forced) s = fma(m3uF, (4.0f / 9.0f) * k, s);
        g[0] = fma(g[0], onem, s);
    }
    const V orw1 = orho * (1.0f / 9.0f), orw2 = orho * (1.0f / 36.0f);
    bgk_pair(g[1], g[3], ux, Fx, m15usq, m3uF, orw1, onem, k, 1.0f / 9.0f, forced);
    bgk_pair(g[2], g[4], uy, Fy, m15usq, m3uF, orw1, onem, k, 1.0f / 9.0f, forced);
    bgk_pair(g[5], g[7], ux + uy, Fx + Fy, m15usq, m3uF, orw2, onem, k, 1.0f / 36.0f, forced);
    bgk_pair(g[6], g[8], uy - ux, Fy - Fx, m15usq, m3uF, orw2, onem, k, 1.0f / 36.0f, forced);
}

// ------------------------------------------------------------------ MRT
// m = M f (rows rho, e, eps, jx, qx, jy, qy, pxx, pxy), m_eq = M f_eq(rho, u) in closed form, force moments
// (compute_forcing_term, MRT.cu:4-27; rows 4 / 5 swapped when QK_D2), f = M^-1 m*.
template <class V>
LBM_HD void collide_mrt_v(const Relax& r, V g[Q], V rho, V ux, V uy, bool forced, V Fx, V Fy) {
    const V a13 = g[1] + g[3], a24 = g[2] + g[4], a57 = g[5] + g[7], a68 = g[6] + g[8];
    const V sA = a13 + a24, sD = a57 + a68;
    V m[Q];
    m[0] = g[0] + (sA + sD);
    m[1] = fma(g[0], -4.0f, fma(sD, 2.0f, -sA));
    m[2] = fma(g[0], 4.0f, fma(sA, -2.0f, sD));
    const V dx1 = g[1] - g[3], dx2 = (g[5] - g[6]) + (g[8] - g[7]);
    const V dy1 = g[2] - g[4], dy2 = (g[5] - g[8]) + (g[6] - g[7]);
    m[3] = dx1 + dx2;
    m[4] = fma(dx1, -2.0f, dx2);
    m[5] = dy1 + dy2;
    m[6] = fma(dy1, -2.0f, dy2);
    m[7] = a13 - a24;
    m[8] = a57 - a68;
    const V jx = rho * ux, jy = rho * uy, usq = fma(ux, ux, uy * uy);
    // d = m_eq - m with m_eq = (rho, rho (3 u^2 - 2), rho (1 - 3 u^2), jx, -jx, jy, -jy, jx ux - jy uy, jx uy).  Every product that
    // feeds the subtraction is written as ONE FMA: ptxas contracts a single-use packed product into a following packed add even
    // though both carry .rn (mul.rn.f32x2 + add.rn.f32x2 -> FFMA2; it never does so for the scalar forms), which made a cell's
    // result depend on whether the one-cell or the two-cell instantiation computed it (tools/v1v2_probe.cu).
    V d[Q];
    d[0] = rho - m[0];
    d[1] = fma(rho, fma(usq, 3.0f, -2.0f), -m[1]);
    d[2] = fma(rho, fma(usq, -3.0f, 1.0f), -m[2]);
    d[3] = fma(rho, ux, -m[3]); d[4] = fma(-rho, ux, -m[4]);
    d[5] = fma(rho, uy, -m[5]); d[6] = fma(-rho, uy, -m[6]);
    d[7] = fma(jx, ux, -(jy * uy)) - m[7];
    d[8] = fma(jx, uy, -m[8]);
LBM_UNROLL
    for (int i = 0; i < Q; i++) m[i] = fma(d[i], r.S[i], m[i]);
    if (forced) {
        const V uF = fma(Fx, ux, Fy * uy);
        V F[Q];
        F[0] = splat<V>(0.0f); F[1] = uF * 6.0f; F[2] = uF * -6.0f; F[3] = Fx;
        if (r.quirks & QK_D2) { F[4] = Fy; F[5] = -Fx; } else { F[4] = -Fx; F[5] = Fy; }
        F[6] = -Fy;
        F[7] = fma(Fx, ux, -(Fy * uy)) * 2.0f;
        F[8] = fma(Fx, uy, Fy * ux);
LBM_UNROLL
        for (int i = 1; i < Q; i++) m[i] = fma(F[i], 1.0f - 0.5f * r.S[i], m[i]);
    }
    const V a = m[0] * (1.0f / 9.0f);
    const V b1 = m[1] * (1.0f / 36.0f), b2 = m[2] * (1.0f / 36.0f);
    const V ax = fma(b2, -2.0f, a - b1), dg = fma(b1, 2.0f, a + b2);
    const V x6 = (m[3] - m[4]) * (1.0f / 6.0f), y6 = (m[5] - m[6]) * (1.0f / 6.0f);
    const V xd = fma(m[4], 1.0f / 12.0f, m[3] * (1.0f / 6.0f)), yd = fma(m[6], 1.0f / 12.0f, m[5] * (1.0f / 6.0f));
    const V p4 = m[7] * 0.25f, q4 = m[8] * 0.25f;
    g[0] = fma(b2 - b1, 4.0f, a);
    const V axp = ax + p4, axm = ax - p4;
    g[1] = axp + x6;
    g[3] = axp - x6;
    g[2] = axm + y6;
    g[4] = axm - y6;
    const V dp = dg + q4, dm = dg - q4, s = xd + yd, t = xd - yd;
    g[5] = dp + s;
    g[7] = dp - s;
    g[6] = dm - t;
    g[8] = dm + t;
}

// ------------------------------------------------------------------ OptimalAdapter
// tau* = theta . (rho/<rho>, rho|u|/<rho|u|>, |Pi|/<|Pi|>, 1), theta = (3e-4, -7.75e-3, 1.6e-4, 8.7e-3); tau* <= 0 (or NaN)
// -> 0.005; clamp 1.5; rate = 1 / (3 tau* + 1/2)  — adapters.cuh:55-109.  The three grid means are the same for every
// cell of a step, so the divisions are multiplications by reciprocals taken once per thread.
struct AdapterAvg { float inv_rho, inv_j, inv_pi; };
LBM_HD float rate_of_tau_star(float t) {
    t = t > 0.0f ? t : 0.005f;          // also taken for NaN (0/0 grid mean of a fluid at rest), as in the reference
    t = fminf(t, 1.5f);
    return fast_rcp(fmaf(3.0f, t, 0.5f));
}
LBM_HD V1 rate_of_tau_star(V1 t) { V1 r; r.a = rate_of_tau_star(t.a); return r; }
LBM_HD V2 rate_of_tau_star(V2 t) { V2 r; r.a = make_float2(rate_of_tau_star(t.a.x), rate_of_tau_star(t.a.y)); return r; }
template <class V>
LBM_HD V optimal_rate_v(V rho, V jmag, V pimag, const AdapterAvg& a) {
    V ts = fma(rho * a.inv_rho, 0.0003f, 0.0087f);
    ts = fma(jmag * a.inv_j, -0.00775f, ts);
    ts = fma(pimag * a.inv_pi, 0.00016f, ts);
    return rate_of_tau_star(ts);
}

// ------------------------------------------------------------------ central moments
// The reference accumulates the nine central moments with per-direction polynomials and multiplies by an 81-entry
// T^-1(u) (cm_matrix_inverse, CM.cuh:141-250, ~900 flop).  Here: raw moments (add chains) -> binomial shift by -u ->
// relax (k_eq = (rho,0,0,2 rho cs2,0,0,0,0,rho cs2^2), force moments (0,Fx,Fy,0,0,0,Fy cs2,Fx cs2,0), CM.cuh:78-99) ->
// shift by +u -> populations.  hi = rate of the moments with index > 5 when the adapter is OptimalAdapter.
template <bool OPTIMAL, class V>
LBM_HD void collide_cm_v(const Relax& r, V g[Q], V ux, V uy, bool forced, V Fx, V Fy, V hi) {
    const V a13 = g[1] + g[3], a24 = g[2] + g[4], a57 = g[5] + g[7], a68 = g[6] + g[8];
    const V d = a57 + a68;
    const V m00 = g[0] + ((a13 + a24) + d);                       // rho recomputed from f (CM.cuh:38-41)
    const V m10 = (g[1] - g[3]) + ((g[5] - g[6]) + (g[8] - g[7]));
    const V m01 = (g[2] - g[4]) + ((g[5] - g[8]) + (g[6] - g[7]));
    const V m20 = a13 + d, m02 = a24 + d;
    const V m11 = a57 - a68;
    const V m21 = (g[5] + g[6]) - (g[7] + g[8]);                   // sum f cx^2 cy
    const V m12 = (g[5] + g[8]) - (g[6] + g[7]);                   // sum f cx cy^2
    const V m22 = d;
    const V ux2 = ux * ux, uy2 = uy * uy, uxuy = ux * uy, tux = ux * 2.0f, tuy = uy * 2.0f;
    // central moments about u
    const V k10 = fma(-ux, m00, m10), k01 = fma(-uy, m00, m01);
    const V k20 = fma(ux2, m00, fma(-tux, m10, m20));
    const V k02 = fma(uy2, m00, fma(-tuy, m01, m02));
    const V k11 = fma(uxuy, m00, fma(-uy, m10, fma(-ux, m01, m11)));
    const V a21 = fma(ux2, m01, fma(-tux, m11, m21));              // sum f (cx-ux)^2 cy
    const V a12 = fma(uy2, m10, fma(-tuy, m11, m12));              // sum f cx (cy-uy)^2
    const V k21 = fma(-uy, k20, a21), k12 = fma(-ux, k02, a12);
    const V k22 = fma(ux2, k02, fma(-tux, a12, fma(uy2, m20, fma(-tuy, m21, m22))));
    const float cs2 = 1.0f / 3.0f;
    V k[Q] = {m00, k10, k01, k20 + k02, k20 - k02, k11, k21, k12, k22};
    const V rho = m00;
    // relaxation towards k_eq: only rows 0, 3 and 8 have a non-zero equilibrium
    k[0] = fma(rho - k[0], r.S[0], k[0]);
    k[1] = fma(-k[1], r.S[1], k[1]);
    k[2] = fma(-k[2], r.S[2], k[2]);
    k[3] = fma(fma(rho, 2.0f * cs2, -k[3]), r.S[3], k[3]);
    k[4] = fma(-k[4], r.S[4], k[4]);
    k[5] = fma(-k[5], r.S[5], k[5]);
    if (OPTIMAL) {
        const V fh = fma(hi, -0.5f, 1.0f);
        k[6] = fma(-k[6], hi, k[6]);
        k[7] = fma(-k[7], hi, k[7]);
        k[8] = fma(fma(rho, cs2 * cs2, -k[8]), hi, k[8]);
        if (forced) {
            k[6] = fma(Fy * cs2, fh, k[6]);
            k[7] = fma(Fx * cs2, fh, k[7]);
        }
    } else {
        k[6] = fma(-k[6], r.S[6], k[6]);
        k[7] = fma(-k[7], r.S[7], k[7]);
        k[8] = fma(fma(rho, cs2 * cs2, -k[8]), r.S[8], k[8]);
        if (forced) {
            k[6] = fma(Fy, cs2 * (1.0f - 0.5f * r.S[6]), k[6]);
            k[7] = fma(Fx, cs2 * (1.0f - 0.5f * r.S[7]), k[7]);
        }
    }
    if (forced) {
        k[1] = fma(Fx, 1.0f - 0.5f * r.S[1], k[1]);
        k[2] = fma(Fy, 1.0f - 0.5f * r.S[2], k[2]);
    }
    // back to raw moments (shift by +u)
    const V c00 = k[0], c10 = k[1], c01 = k[2];
    const V c20 = (k[3] + k[4]) * 0.5f, c02 = (k[3] - k[4]) * 0.5f;
    const V c11 = k[5], c21 = k[6], c12 = k[7], c22 = k[8];
    const V r10 = fma(ux, c00, c10), r01 = fma(uy, c00, c01);
    const V r20 = fma(ux2, c00, fma(tux, c10, c20));
    const V r02 = fma(uy2, c00, fma(tuy, c01, c02));
    const V r11 = fma(uxuy, c00, fma(uy, c10, fma(ux, c01, c11)));
    const V b12 = fma(uy2, c10, fma(tuy, c11, c12));               // sum f (cx-ux) cy^2
    const V r21 = fma(uy, r20, fma(ux2, c01, fma(tux, c11, c21)));
    const V r12 = fma(ux, r02, b12);
    const V r22 = fma(ux2, r02, fma(tux, b12, fma(uy2, c20, fma(tuy, c21, c22))));
    g[0] = (c00 - r20) + (r22 - r02);
    const V e1 = r20 - r22, o1 = r10 - r12;                        // g1 + g3 = r20 - r22,  g1 - g3 = r10 - r12
    g[1] = (e1 + o1) * 0.5f;
    g[3] = (e1 - o1) * 0.5f;
    const V e2 = r02 - r22, o2 = r01 - r21;
    g[2] = (e2 + o2) * 0.5f;
    g[4] = (e2 - o2) * 0.5f;
    const V s1 = r11 + r22, s2 = r22 - r11, t1 = r21 + r12, t2 = r21 - r12;
    g[5] = (s1 + t1) * 0.25f;
    g[7] = (s1 - t1) * 0.25f;
    g[6] = (s2 + t2) * 0.25f;
    g[8] = (s2 - t2) * 0.25f;
}

}  // namespace lbm
