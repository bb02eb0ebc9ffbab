// CPU check of cuda_lbm_b200/csrc/collide.cuh: every collision operator, instantiated for V1 (one cell) and V2 (two packed
// cells) and executed ON THE HOST, against a direct fp64 evaluation of the operator's definition (dense 9x9 matrices, generic
// linear solves) — an independent restatement of the algebra, not of the code.  Also: V2 lanes == V1 bit for bit.
// Built and run by tests/test_host_math.py (nvcc compiles it as host code; no GPU involved).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include "../cuda_lbm_b200/csrc/collide.cuh"

using namespace lbm;
static const int CX[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1}, CY[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
static const double W[9] = {4. / 9, 1. / 9, 1. / 9, 1. / 9, 1. / 9, 1. / 36, 1. / 36, 1. / 36, 1. / 36};

static double urand() { return rand() / (RAND_MAX + 1.0); }

static void solve9(double A[9][9], double b[9], double x[9]) {      // Gaussian elimination with partial pivoting
    double M[9][10];
    for (int i = 0; i < 9; i++) { for (int j = 0; j < 9; j++) M[i][j] = A[i][j]; M[i][9] = b[i]; }
    for (int c = 0; c < 9; c++) {
        int p = c;
        for (int r = c + 1; r < 9; r++) if (fabs(M[r][c]) > fabs(M[p][c])) p = r;
        for (int j = 0; j < 10; j++) std::swap(M[c][j], M[p][j]);
        for (int r = 0; r < 9; r++) if (r != c) { double f = M[r][c] / M[c][c]; for (int j = c; j < 10; j++) M[r][j] -= f * M[c][j]; }
    }
    for (int i = 0; i < 9; i++) x[i] = M[i][9] / M[i][i];
}

static void feq64(double rho, double ux, double uy, double fe[9]) {
    for (int q = 0; q < 9; q++) { double cu = CX[q] * ux + CY[q] * uy; fe[q] = W[q] * rho * (1 + 3 * cu + 4.5 * cu * cu - 1.5 * (ux * ux + uy * uy)); }
}
static void guo64(double ux, double uy, double Fx, double Fy, double Fh[9]) {     // w [(c-u)/cs2 + (c.u) c/cs4] . F
    for (int q = 0; q < 9; q++) {
        double cu = CX[q] * ux + CY[q] * uy;
        Fh[q] = W[q] * (3 * ((CX[q] - ux) * Fx + (CY[q] - uy) * Fy) + 9 * cu * (CX[q] * Fx + CY[q] * Fy));
    }
}
static void mrt_rows(double M[9][9]) {          // Lallemand-Luo basis: rho, e, eps, jx, qx, jy, qy, pxx, pxy
    for (int q = 0; q < 9; q++) {
        double x = CX[q], y = CY[q], c2 = x * x + y * y;
        M[0][q] = 1; M[1][q] = -4 + 3 * c2; M[2][q] = 4 - 10.5 * c2 + 4.5 * c2 * c2; M[3][q] = x; M[4][q] = (-5 + 3 * c2) * x;
        M[5][q] = y; M[6][q] = (-5 + 3 * c2) * y; M[7][q] = x * x - y * y; M[8][q] = x * y;
    }
}
static void cm_rows(double ux, double uy, double T[9][9]) {     // central-moment basis about u (CM.cuh:51-67)
    for (int q = 0; q < 9; q++) {
        double x = CX[q] - ux, y = CY[q] - uy;
        T[0][q] = 1; T[1][q] = x; T[2][q] = y; T[3][q] = x * x + y * y; T[4][q] = x * x - y * y; T[5][q] = x * y;
        T[6][q] = x * x * y; T[7][q] = x * y * y; T[8][q] = x * x * y * y;
    }
}

struct Case { double g[9], rho, ux, uy, Fx, Fy; };

static void ref_bgk(const Case& c, const Relax& r, double out[9]) {
    double fe[9], Fh[9];
    feq64(c.rho, c.ux, c.uy, fe); guo64(c.ux, c.uy, c.Fx, c.Fy, Fh);
    for (int q = 0; q < 9; q++) out[q] = c.g[q] - r.omega * (c.g[q] - fe[q]) + (1 - 0.5 * r.omega) * Fh[q];
}
static void ref_mrt(const Case& c, const Relax& r, double out[9]) {
    double M[9][9], fe[9], Fh[9], m[9], me[9], mF[9];
    mrt_rows(M); feq64(c.rho, c.ux, c.uy, fe); guo64(c.ux, c.uy, c.Fx, c.Fy, Fh);
    for (int i = 0; i < 9; i++) { m[i] = me[i] = mF[i] = 0; for (int q = 0; q < 9; q++) { m[i] += M[i][q] * c.g[q]; me[i] += M[i][q] * fe[q]; mF[i] += M[i][q] * Fh[q]; } }
    if (r.quirks & QK_D2) { mF[4] = c.Fy; mF[5] = -c.Fx; }          // the reference's row order
    for (int i = 0; i < 9; i++) m[i] = m[i] - r.S[i] * (m[i] - me[i]) + (1 - 0.5 * r.S[i]) * mF[i];
    solve9(M, m, out);
}
static void ref_cm(const Case& c, const Relax& r, bool optimal, double hi, double out[9]) {
    double T[9][9], k[9];
    cm_rows(c.ux, c.uy, T);
    for (int i = 0; i < 9; i++) { k[i] = 0; for (int q = 0; q < 9; q++) k[i] += T[i][q] * c.g[q]; }
    double rho = 0; for (int q = 0; q < 9; q++) rho += c.g[q];
    const double cs2 = 1. / 3;
    double keq[9] = {rho, 0, 0, 2 * rho * cs2, 0, 0, 0, 0, rho * cs2 * cs2}, F[9] = {0, c.Fx, c.Fy, 0, 0, 0, c.Fy * cs2, c.Fx * cs2, 0};
    for (int i = 0; i < 9; i++) { double s = (optimal && i > 5) ? hi : r.S[i]; k[i] = k[i] - s * (k[i] - keq[i]) + (1 - 0.5 * s) * F[i]; }
    solve9(T, k, out);
}

template <class V> static void load(V g[9], const Case* c, int lane);
template <> void load<V1>(V1 g[9], const Case* c, int) { for (int q = 0; q < 9; q++) g[q].a = (float)c[0].g[q]; }
template <> void load<V2>(V2 g[9], const Case* c, int) { for (int q = 0; q < 9; q++) g[q].a = make_float2((float)c[0].g[q], (float)c[1].g[q]); }
static V1 mk1(const Case* c, double Case::*f) { return V1{(float)(c[0].*f)}; }
static V2 mk2(const Case* c, double Case::*f) { return V2{make_float2((float)(c[0].*f), (float)(c[1].*f))}; }

int main() {
    srand(12345);
    int bad = 0;
    double worst[4] = {0, 0, 0, 0};
    for (int trial = 0; trial < 4000; trial++) {
        Case c[2];
        Relax r;
        const bool forced = trial % 3 != 0;
        r.omega = (float)(0.55 + 1.4 * urand());
        r.quirks = (trial & 1) ? 63 : 0;
        for (int i = 0; i < 9; i++) r.S[i] = (float)(2.0 * urand());
        if (trial % 5 == 0) { float S0[9] = {0, r.omega, r.omega, 0, r.omega, 0, r.omega, r.omega, r.omega}; memcpy(r.S, S0, sizeof(S0)); }
        for (int l = 0; l < 2; l++) {
            double rho0 = 0.8 + 0.4 * urand(), u0 = 0.2 * (urand() - 0.5), v0 = 0.2 * (urand() - 0.5), fe[9];
            feq64(rho0, u0, v0, fe);
            for (int q = 0; q < 9; q++) c[l].g[q] = (double)(float)(fe[q] * (1 + 0.1 * (urand() - 0.5)));
            c[l].Fx = forced ? (double)(float)(1e-3 * (urand() - 0.5)) : 0.0;
            c[l].Fy = forced ? (double)(float)(1e-3 * (urand() - 0.5)) : 0.0;
        }
        // moments through the template (fp32) — also the inputs of the collision, as in the kernels
        V1 g1[2][9]; V2 g2[9];
        load<V1>(g1[0], &c[0], 0); load<V1>(g1[1], &c[1], 0); load<V2>(g2, c, 0);
        Mom<V1> m1[2] = {moments_v(g1[0]), moments_v(g1[1])};
        Mom<V2> m2 = moments_v(g2);
        for (int l = 0; l < 2; l++) {
            double rho = 0, jx = 0, jy = 0, pxx = 0, pxy = 0, pyy = 0;
            for (int q = 0; q < 9; q++) { rho += c[l].g[q]; jx += c[l].g[q] * CX[q]; jy += c[l].g[q] * CY[q]; pxx += c[l].g[q] * CX[q] * CX[q]; pxy += c[l].g[q] * CX[q] * CY[q]; pyy += c[l].g[q] * CY[q] * CY[q]; }
            if (fabs(m1[l].rho.a - rho) > 3e-7 || fabs(m1[l].ux.a - jx / rho) > 3e-7 || fabs(m1[l].uy.a - jy / rho) > 3e-7 || fabs(m1[l].pxx.a - pxx) > 3e-7 ||
                fabs(m1[l].pxy.a - pxy) > 3e-7 || fabs(m1[l].pyy.a - pyy) > 3e-7) { printf("moments mismatch trial %d\n", trial); bad++; }
            double pin = sqrt(pxx * pxx + 2 * pxy * pxy + pyy * pyy);
            if (fabs(pi_norm_v(m1[l]).a - pin) > 5e-7) { printf("pi_norm mismatch trial %d\n", trial); bad++; }
            // u corrected by F/2rho as the kernels do, in fp32, then handed to both sides
            float h = m1[l].inv_rho.a * 0.5f;
            c[l].rho = m1[l].rho.a;
            c[l].ux = forced ? fmaf((float)c[l].Fx, h, m1[l].ux.a) : m1[l].ux.a;
            c[l].uy = forced ? fmaf((float)c[l].Fy, h, m1[l].uy.a) : m1[l].uy.a;
        }
        const float lanes2[2] = {m2.rho.a.x, m2.rho.a.y};
        if (lanes2[0] != m1[0].rho.a || lanes2[1] != m1[1].rho.a || m2.ux.a.x != m1[0].ux.a || m2.uy.a.y != m1[1].uy.a) { printf("V2 moments != V1 trial %d\n", trial); bad++; }
        const float hi[2] = {(float)(1.8 + 0.19 * urand()), (float)(1.8 + 0.19 * urand())};
        for (int op = 0; op < 4; op++) {
            double ref[2][9];
            for (int l = 0; l < 2; l++) {
                if (op == 0) ref_bgk(c[l], r, ref[l]); else if (op == 1) ref_mrt(c[l], r, ref[l]);
                else ref_cm(c[l], r, op == 3, hi[l], ref[l]);
            }
            V1 a[2][9]; V2 b[9];
            load<V1>(a[0], &c[0], 0); load<V1>(a[1], &c[1], 0); load<V2>(b, c, 0);
            for (int l = 0; l < 2; l++) {
                const V1 rho = mk1(&c[l], &Case::rho), ux = mk1(&c[l], &Case::ux), uy = mk1(&c[l], &Case::uy), Fx = mk1(&c[l], &Case::Fx), Fy = mk1(&c[l], &Case::Fy);
                if (op == 0) collide_bgk_v(r, a[l], rho, ux, uy, forced, Fx, Fy);
                else if (op == 1) collide_mrt_v(r, a[l], rho, ux, uy, forced, Fx, Fy);
                else if (op == 2) collide_cm_v<false>(r, a[l], ux, uy, forced, Fx, Fy, V1{1.0f});
                else collide_cm_v<true>(r, a[l], ux, uy, forced, Fx, Fy, V1{hi[l]});
            }
            {
                const V2 rho = mk2(c, &Case::rho), ux = mk2(c, &Case::ux), uy = mk2(c, &Case::uy), Fx = mk2(c, &Case::Fx), Fy = mk2(c, &Case::Fy);
                if (op == 0) collide_bgk_v(r, b, rho, ux, uy, forced, Fx, Fy);
                else if (op == 1) collide_mrt_v(r, b, rho, ux, uy, forced, Fx, Fy);
                else if (op == 2) collide_cm_v<false>(r, b, ux, uy, forced, Fx, Fy, splat<V2>(1.0f));
                else collide_cm_v<true>(r, b, ux, uy, forced, Fx, Fy, V2{make_float2(hi[0], hi[1])});
            }
            for (int q = 0; q < 9; q++) {
                if (b[q].a.x != a[0][q].a || b[q].a.y != a[1][q].a) { printf("op %d trial %d q %d: V2 lanes differ from V1\n", op, trial, q); bad++; }
                for (int l = 0; l < 2; l++) {
                    double d = fabs(a[l][q].a - ref[l][q]);
                    if (d > worst[op]) worst[op] = d;
                    if (!(d <= 1.5e-6)) { if (bad < 20) printf("op %d trial %d lane %d q %d: %.9g vs %.9g (forced %d quirks %d)\n", op, trial, l, q, a[l][q].a, ref[l][q], forced, r.quirks); bad++; }
                }
            }
        }
    }
    // adapter rate
    AdapterAvg av{1.0f / 1.01f, 1.0f / 0.02f, 1.0f / 0.6f};
    for (int i = 0; i < 1000; i++) {
        float rho = (float)(0.9 + 0.2 * urand()), j = (float)(0.1 * urand()), pi = (float)(0.5 + 0.3 * urand());
        double ts = 0.0003 * rho / 1.01 - 0.00775 * j / 0.02 + 0.00016 * pi / 0.6 + 0.0087;
        if (!(ts > 0)) ts = 0.005;
        if (ts > 1.5) ts = 1.5;
        double want = 1 / (3 * ts + 0.5);
        float got = optimal_rate_v(V1{rho}, V1{j}, V1{pi}, av).a;
        V2 got2 = optimal_rate_v(V2{make_float2(rho, rho)}, V2{make_float2(j, j)}, V2{make_float2(pi, pi)}, av);
        if (fabs(got - want) > 2e-5 * want || got2.a.x != got || got2.a.y != got) { printf("adapter rate mismatch %g vs %g\n", got, want); bad++; }
    }
    float nanr = optimal_rate_v(V1{1.0f}, V1{NAN}, V1{0.5f}, av).a;        // 0/0 grid mean at t = 1 of a cavity: tau* -> 0.005
    if (fabs(nanr - 1.0f / (3 * 0.005f + 0.5f)) > 1e-6) { printf("adapter NaN handling: %g\n", nanr); bad++; }
    printf("worst |fp32 - fp64 definition|: BGK %.2e MRT %.2e CM %.2e CM_OPT %.2e\n", worst[0], worst[1], worst[2], worst[3]);
    printf(bad ? "FAILED %d\n" : "OK\n", bad);
    return bad ? 1 : 0;
}
