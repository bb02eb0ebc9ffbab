#!/usr/bin/env python
"""Offline derivation and proof of the D2Q9 moment transforms used by cuda_lbm_b200/csrc/collide.cuh.

The reference generates two pieces of its hot path with sympy (src/codegen/gen.py -> f_eq of equilibrium.cu:5-39,
src/codegen/cm_matrix_inv.py -> the 81 polynomial entries of T^-1(u) in CM.cuh:141-250) and multiplies dense 9x9 matrices
from constant memory for MRT (MRT.cu:29-76).  The B200 kernels evaluate the same linear maps as folded add/FMA chains:

  MRT   m = M f and f = M^-1 m with the 0 / +-1 / +-2 / +-4 entries resolved and opposite directions sharing sub-sums
  CM    raw moments (integer chains) -> binomial shift by -u -> relax -> shift by +u -> populations, instead of T^-1(u) k
  BGK   opposite directions share the symmetric part of f_eq and of the Guo force term

This script states the reference's definitions (file:line cited), transcribes the chains of collide.cuh statement by statement
with the same names, and proves with exact rational arithmetic that both are the same polynomials:  `verify()` raises on the
first mismatch (tests/test_codegen_identities.py runs it), `--emit` prints the folded chains as C text the way a generator
would.  The compiled code itself is checked numerically against a direct fp64 evaluation by tests/host_math_check.cu.
"""
import argparse
import sys

import sympy as sp

Q = 9
# h_C, h_OPP, h_weights — reference src/core/lbm_constants.cuh:13-31
CX = [0, 1, 0, -1, 0, 1, -1, -1, 1]
CY = [0, 0, 1, 0, -1, 1, 1, -1, -1]
W = [sp.Rational(4, 9)] + [sp.Rational(1, 9)] * 4 + [sp.Rational(1, 36)] * 4
OPP = [0, 3, 4, 1, 2, 7, 8, 5, 6]

f = sp.symbols("f0:9")
ux, uy, rho, Fx, Fy, omega = sp.symbols("ux uy rho Fx Fy omega")
S = sp.symbols("S0:9")


def fma(a, b, c):
    return a * b + c


def same(a, b, what):
    d = sp.expand(sp.sympify(a) - sp.sympify(b))
    if d != 0:
        raise AssertionError(f"{what}: differs by {d}")


# ------------------------------------------------------------------ definitions of the reference
def mrt_matrix():
    """h_M rows rho, e, eps, jx, qx, jy, qy, pxx, pxy — src/core/lbm_constants.cuh:33-43 (Lallemand & Luo basis)."""
    rows = []
    for q in range(Q):
        cx, cy = CX[q], CY[q]
        c2 = cx * cx + cy * cy
        rows.append([1, -4 + 3 * c2, 4 - sp.Rational(21, 2) * c2 + sp.Rational(9, 2) * c2 * c2, cx, (-5 + 3 * c2) * cx, cy, (-5 + 3 * c2) * cy,
                     cx * cx - cy * cy, cx * cy])
    return sp.Matrix(rows).T


def cm_matrix():
    """T(u): central moments k_0..8 about u — src/core/collision/CM/CM.cuh:53-72 (and src/codegen/cm_matrix_inv.py's M)."""
    T = sp.zeros(Q, Q)
    for q in range(Q):
        a, b = CX[q] - ux, CY[q] - uy
        col = [1, a, b, a * a + b * b, a * a - b * b, a * b, a * a * b, a * b * b, a * a * b * b]
        for i in range(Q):
            T[i, q] = col[i]
    return T


def feq(q, r, vx, vy):
    """equilibrium_node — src/core/equilibrium/equilibrium.cu:5-39 (what src/codegen/gen.py emits): second order in u."""
    cu = CX[q] * vx + CY[q] * vy
    return W[q] * r * (1 + 3 * cu + sp.Rational(9, 2) * cu * cu - sp.Rational(3, 2) * (vx * vx + vy * vy))


def guo(q, vx, vy, gx, gy, om):
    """BGK<2>::apply force term — src/core/collision/BGK/BGK.cuh:34-48: w (1 - omega/2) [ (c-u)/cs2 + (c.u) c / cs4 ] . F"""
    cu = CX[q] * vx + CY[q] * vy
    return W[q] * (1 - sp.Rational(1, 2) * om) * (3 * ((CX[q] - vx) * gx + (CY[q] - vy) * gy) + 9 * cu * (CX[q] * gx + CY[q] * gy))


# ------------------------------------------------------------------ the chains of collide.cuh, transcribed
def moments_chain(g):
    """moments_v"""
    r = (((((((g[0] + g[1]) + g[2]) + g[3]) + g[4]) + g[5]) + g[6]) + g[7]) + g[8]
    d56, d87, d58, d67 = g[5] - g[6], g[8] - g[7], g[5] - g[8], g[6] - g[7]
    jx = ((g[1] - g[3]) + d56) + d87
    jy = ((g[2] - g[4]) + d58) + d67
    d = ((g[5] + g[6]) + g[7]) + g[8]
    return r, jx, jy, (g[1] + g[3]) + d, d56 - d87, (g[2] + g[4]) + d


def mrt_forward_chain(g):
    """collide_mrt_v, m = M f"""
    a13, a24, a57, a68 = g[1] + g[3], g[2] + g[4], g[5] + g[7], g[6] + g[8]
    sA, sD = a13 + a24, a57 + a68
    m = [None] * Q
    m[0] = g[0] + (sA + sD)
    m[1] = fma(g[0], -4, fma(sD, 2, -sA))
    m[2] = fma(g[0], 4, fma(sA, -2, sD))
    dx1, dx2 = g[1] - g[3], (g[5] - g[6]) + (g[8] - g[7])
    dy1, dy2 = g[2] - g[4], (g[5] - g[8]) + (g[6] - g[7])
    m[3] = dx1 + dx2
    m[4] = fma(dx1, -2, dx2)
    m[5] = dy1 + dy2
    m[6] = fma(dy1, -2, dy2)
    m[7] = a13 - a24
    m[8] = a57 - a68
    return m


def mrt_equilibrium_chain(r, vx, vy):
    """collide_mrt_v, m_eq = M f_eq(rho, u) in closed form"""
    jx, jy, usq = r * vx, r * vy, fma(vx, vx, vy * vy)
    return [r, r * fma(usq, 3, -2), r * fma(usq, -3, 1), jx, -jx, jy, -jy, fma(jx, vx, -(jy * vy)), jx * vy]


def mrt_backward_chain(m):
    """collide_mrt_v, f = M^-1 m"""
    g = [None] * Q
    a = m[0] * sp.Rational(1, 9)
    b1, b2 = m[1] * sp.Rational(1, 36), m[2] * sp.Rational(1, 36)
    ax, dg = fma(b2, -2, a - b1), fma(b1, 2, a + b2)
    x6, y6 = (m[3] - m[4]) * sp.Rational(1, 6), (m[5] - m[6]) * sp.Rational(1, 6)
    xd, yd = fma(m[4], sp.Rational(1, 12), m[3] * sp.Rational(1, 6)), fma(m[6], sp.Rational(1, 12), m[5] * sp.Rational(1, 6))
    p4, q4 = m[7] * sp.Rational(1, 4), m[8] * sp.Rational(1, 4)
    g[0] = fma(b2 - b1, 4, a)
    axp, axm = ax + p4, ax - p4
    g[1], g[3] = axp + x6, axp - x6
    g[2], g[4] = axm + y6, axm - y6
    dp, dm, s, t = dg + q4, dg - q4, xd + yd, xd - yd
    g[5], g[7] = dp + s, dp - s
    g[6], g[8] = dm - t, dm + t
    return g


def mrt_force_chain(vx, vy, gx, gy, swapped):
    """collide_mrt_v, force moments (compute_forcing_term, MRT.cu:4-27); swapped = the reference's row order (A-D2)"""
    uF = fma(gx, vx, gy * vy)
    F = [0, uF * 6, uF * -6, gx, None, None, -gy, fma(gx, vx, -(gy * vy)) * 2, fma(gx, vy, gy * vx)]
    F[4], F[5] = (gy, -gx) if swapped else (-gx, gy)
    return F


def cm_forward_chain(g, vx, vy):
    """collide_cm_v: raw moments -> central moments k_0..8"""
    a13, a24, a57, a68 = g[1] + g[3], g[2] + g[4], g[5] + g[7], g[6] + g[8]
    d = a57 + a68
    m00 = g[0] + ((a13 + a24) + d)
    m10 = (g[1] - g[3]) + ((g[5] - g[6]) + (g[8] - g[7]))
    m01 = (g[2] - g[4]) + ((g[5] - g[8]) + (g[6] - g[7]))
    m20, m02 = a13 + d, a24 + d
    m11 = a57 - a68
    m21 = (g[5] + g[6]) - (g[7] + g[8])
    m12 = (g[5] + g[8]) - (g[6] + g[7])
    m22 = d
    ux2, uy2, uxuy, tux, tuy = vx * vx, vy * vy, vx * vy, vx * 2, vy * 2
    k10, k01 = fma(-vx, m00, m10), fma(-vy, m00, m01)
    k20 = fma(ux2, m00, fma(-tux, m10, m20))
    k02 = fma(uy2, m00, fma(-tuy, m01, m02))
    k11 = fma(uxuy, m00, fma(-vy, m10, fma(-vx, m01, m11)))
    a21 = fma(ux2, m01, fma(-tux, m11, m21))
    a12 = fma(uy2, m10, fma(-tuy, m11, m12))
    k21, k12 = fma(-vy, k20, a21), fma(-vx, k02, a12)
    k22 = fma(ux2, k02, fma(-tux, a12, fma(uy2, m20, fma(-tuy, m21, m22))))
    return [m00, k10, k01, k20 + k02, k20 - k02, k11, k21, k12, k22]


def cm_backward_chain(k, vx, vy):
    """collide_cm_v: central moments -> raw moments (shift by +u) -> populations; stands for T^-1(u) k (CM.cuh:121-128,141-250)"""
    ux2, uy2, uxuy, tux, tuy = vx * vx, vy * vy, vx * vy, vx * 2, vy * 2
    c00, c10, c01 = k[0], k[1], k[2]
    c20, c02 = (k[3] + k[4]) * sp.Rational(1, 2), (k[3] - k[4]) * sp.Rational(1, 2)
    c11, c21, c12, c22 = k[5], k[6], k[7], k[8]
    r10, r01 = fma(vx, c00, c10), fma(vy, c00, c01)
    r20 = fma(ux2, c00, fma(tux, c10, c20))
    r02 = fma(uy2, c00, fma(tuy, c01, c02))
    r11 = fma(uxuy, c00, fma(vy, c10, fma(vx, c01, c11)))
    b12 = fma(uy2, c10, fma(tuy, c11, c12))
    r21 = fma(vy, r20, fma(ux2, c01, fma(tux, c11, c21)))
    r12 = fma(vx, r02, b12)
    r22 = fma(ux2, r02, fma(tux, b12, fma(uy2, c20, fma(tuy, c21, c22))))
    g = [None] * Q
    g[0] = (c00 - r20) + (r22 - r02)
    e1, o1 = r20 - r22, r10 - r12
    g[1], g[3] = (e1 + o1) / 2, (e1 - o1) / 2
    e2, o2 = r02 - r22, r01 - r21
    g[2], g[4] = (e2 + o2) / 2, (e2 - o2) / 2
    s1, s2, t1, t2 = r11 + r22, r22 - r11, r21 + r12, r21 - r12
    g[5], g[7] = (s1 + t1) / 4, (s1 - t1) / 4
    g[6], g[8] = (s2 + t2) / 4, (s2 - t2) / 4
    return g


def bgk_chain(g, r, vx, vy, gx, gy, om, forced):
    """collide_bgk_v / bgk_pair"""
    onem, k = 1 - om, 1 - sp.Rational(1, 2) * om
    m15usq = fma(vx, vx, vy * vy) * sp.Rational(-3, 2)
    orho = r * om
    m3uF = fma(vx, gx, vy * gy) * -3 if forced else 0
    out = list(g)
    orw0 = orho * sp.Rational(4, 9)
    s = fma(orw0, m15usq, orw0)
    if forced:
        s = fma(m3uF, sp.Rational(4, 9) * k, s)
    out[0] = fma(g[0], onem, s)

    def pair(qa, qb, cu, cF, orw, w):
        sym = fma(orw, fma(cu * sp.Rational(9, 2), cu, m15usq), orw)
        anti = (orw * 3) * cu
        if forced:
            sym = fma(fma(cu * 9, cF, m3uF), w * k, sym)
            anti = fma(cF, 3 * w * k, anti)
        out[qa] = fma(g[qa], onem, sym + anti)
        out[qb] = fma(g[qb], onem, sym - anti)

    orw1, orw2 = orho * sp.Rational(1, 9), orho * sp.Rational(1, 36)
    pair(1, 3, vx, gx, orw1, sp.Rational(1, 9))
    pair(2, 4, vy, gy, orw1, sp.Rational(1, 9))
    pair(5, 7, vx + vy, gx + gy, orw2, sp.Rational(1, 36))
    pair(6, 8, vy - vx, gy - gx, orw2, sp.Rational(1, 36))
    return out


# ------------------------------------------------------------------ the proofs
def verify(verbose=False):
    checks = 0
    fv = sp.Matrix(f)
    # lattice identities
    assert sum(W) == 1 and all(OPP[OPP[q]] == q and CX[OPP[q]] == -CX[q] and CY[OPP[q]] == -CY[q] for q in range(Q))
    # moments
    r, jx, jy, pxx, pxy, pyy = moments_chain(f)
    for got, want, name in ((r, sum(f), "rho"), (jx, sum(f[q] * CX[q] for q in range(Q)), "jx"), (jy, sum(f[q] * CY[q] for q in range(Q)), "jy"),
                            (pxx, sum(f[q] * CX[q] ** 2 for q in range(Q)), "Pxx"), (pxy, sum(f[q] * CX[q] * CY[q] for q in range(Q)), "Pxy"),
                            (pyy, sum(f[q] * CY[q] ** 2 for q in range(Q)), "Pyy")):
        same(got, want, "moments " + name); checks += 1
    # MRT
    M = mrt_matrix()
    Minv = M.inv()
    assert M * Minv == sp.eye(Q)
    for i, (got, want) in enumerate(zip(mrt_forward_chain(f), M * fv)):
        same(got, want, f"MRT m[{i}] = (M f)[{i}]"); checks += 1
    m = sp.symbols("m0:9")
    for i, (got, want) in enumerate(zip(mrt_backward_chain(m), Minv * sp.Matrix(m))):
        same(got, want, f"MRT g[{i}] = (M^-1 m)[{i}]"); checks += 1
    feq_vec = sp.Matrix([feq(q, rho, ux, uy) for q in range(Q)])
    for i, (got, want) in enumerate(zip(mrt_equilibrium_chain(rho, ux, uy), M * feq_vec)):
        same(got, want, f"MRT m_eq[{i}] = (M f_eq)[{i}]"); checks += 1
    # force moments: with the rows in M's order (quirk D2 repaired) they are M applied to the Guo term at omega = 0,
    # i.e. the (1 - S/2)-weighted source of Guo's scheme in moment space
    guo0 = sp.Matrix([guo(q, ux, uy, Fx, Fy, 0) for q in range(Q)])
    for i, (got, want) in enumerate(zip(mrt_force_chain(ux, uy, Fx, Fy, swapped=False), M * guo0)):
        same(got, want, f"MRT F[{i}] = (M guo)[{i}]"); checks += 1
    sw = mrt_force_chain(ux, uy, Fx, Fy, swapped=True)
    same(sw[4], Fy, "MRT F[4] with the reference's row order (A-D2)"); same(sw[5], -Fx, "MRT F[5] with the reference's row order (A-D2)"); checks += 2
    # BGK == MRT with S = omega on every row (scenario.cuh:42-57), without and with force
    for forced in (False, True):
        b = bgk_chain(f, sum(f), ux, uy, Fx, Fy, omega, forced)
        for q in range(Q):
            want = f[q] - omega * (f[q] - feq(q, sum(f), ux, uy)) + (guo(q, ux, uy, Fx, Fy, omega) if forced else 0)
            same(b[q], want, f"BGK f'[{q}] forced={forced}"); checks += 1
        mm = mrt_forward_chain(f)
        me = mrt_equilibrium_chain(sum(f), ux, uy)
        mm = [fma(me[i] - mm[i], omega, mm[i]) for i in range(Q)]
        if forced:
            F = mrt_force_chain(ux, uy, Fx, Fy, swapped=False)
            mm = [mm[0]] + [fma(F[i], 1 - sp.Rational(1, 2) * omega, mm[i]) for i in range(1, Q)]
        for q, (got, want) in enumerate(zip(mrt_backward_chain(mm), b)):
            same(got, want, f"MRT(S=omega) == BGK, f'[{q}] forced={forced}"); checks += 1
    # CM: the shift chains are T(u) and T(u)^-1
    T = cm_matrix()
    for i, (got, want) in enumerate(zip(cm_forward_chain(f, ux, uy), T * fv)):
        same(got, want, f"CM k[{i}] = (T(u) f)[{i}]"); checks += 1
    k = sp.symbols("k0:9")
    back = cm_backward_chain(k, ux, uy)
    for i, (got, want) in enumerate(zip(T * sp.Matrix(back), k)):       # T(u) . chain(k) == k  <=>  chain == T(u)^-1
        same(got, want, f"CM T(u) T^-1(u) row {i}"); checks += 1
    # k_eq (CM.cuh:78-86) is T(u) f_eq up to the O(u^2)-in-cs-moments truncation the reference makes: rows 0..5 exactly
    keq = T * feq_vec
    want = [rho, 0, 0, sp.Rational(2, 3) * rho, 0, 0]
    for i in range(6):
        same(keq[i], want[i], f"CM k_eq[{i}]"); checks += 1
    # CM at u = 0 is a raw-moment relaxation: T(0) has integer entries only
    assert all(v.is_integer for v in T.subs({ux: 0, uy: 0}))
    if verbose:
        print(f"{checks} polynomial identities hold")
    return checks


def emit(out=sys.stdout):
    """The folded chains as C text (float literals), common sub-expressions named — what a generator would hand to collide.cuh."""
    from sympy.printing.c import C99CodePrinter

    class P(C99CodePrinter):
        def _print_Rational(self, e):
            return f"{float(e)!r}f" if e.q != 1 else f"{e.p}.0f"

        def _print_Integer(self, e):
            return f"{int(e)}.0f"

        def _print_Pow(self, e):
            if e.exp.is_Integer and 0 < int(e.exp) <= 4:
                return "(" + "*".join([self._print(e.base)] * int(e.exp)) + ")"
            return super()._print_Pow(e)

    pr = P()
    g = sp.symbols("g0:9")
    blocks = [("moments (macroscopics.cu:5-38)", ["rho", "jx", "jy", "Pxx", "Pxy", "Pyy"], list(moments_chain(g))),
              ("m = M f (MRT.cu:29-45)", [f"m[{i}]" for i in range(Q)], mrt_forward_chain(g)),
              ("f = M^-1 m (MRT.cu:60-76)", [f"g[{i}]" for i in range(Q)], mrt_backward_chain(sp.symbols("m0:9"))),
              ("k = T(u) f (CM.cuh:53-72)", [f"k[{i}]" for i in range(Q)], cm_forward_chain(g, ux, uy)),
              ("f = T^-1(u) k (CM.cuh:121-128,141-250)", [f"g[{i}]" for i in range(Q)], cm_backward_chain(sp.symbols("k0:9"), ux, uy)),
              ("f_eq (equilibrium.cu:5-39)", [f"feq[{q}]" for q in range(Q)], [feq(q, rho, ux, uy) for q in range(Q)])]
    for title, names, exprs in blocks:
        print(f"// {title}", file=out)
        repl, red = sp.cse([sp.expand(e) for e in exprs], optimizations="basic")
        for s, e in repl:
            print(f"const float {s} = {pr.doprint(e)};", file=out)
        for n, e in zip(names, red):
            print(f"{n} = {pr.doprint(e)};", file=out)
        print(file=out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--emit", action="store_true", help="print the folded chains as C text")
    a = ap.parse_args()
    verify(verbose=True)
    if a.emit:
        emit()
