/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * CPU restatement (plain C, fp32 with the reference's own fp64 islands) of the D2Q9 hot
 * path of Carabalone/cuda-lbm, kernel by kernel, in the reference's AoS layout
 * (f[node*9+q], u[node*2+c]; the only layout in which its 2-D path is self-consistent,
 * SURVEY.md A-D4).  All citations are file:line under /root/reference/src/.
 *
 * Parity status: the reference ships NO tests / golden vectors.  This
 * restatement is pinned instead against outputs of the reference's own CUDA solver run on a
 * B200 (oracle/build_ref.sh + oracle/ref_cuda/ref_driver.cu -> tests/golden/*.npz, checked
 * by tests/test_oracle_golden.py).  Because the reference executes on a GPU (FMA
 * contraction chosen by nvcc) and this file on a CPU (-ffp-contract=off), agreement is to
 * fp32 round-off, not bit-exact; the tolerances are stated in the tests.
 *
 * Every reference defect on the path is reproduced when its quirk
 * bit is set (default: all set == "what the reference computes") and repaired when clear.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define Q 9

/* quirk bits — 1 = reproduce the reference */
#define QK_D1_STALE_F0     1   /* streaming.cu:9   rest population never copied to f_back        */
#define QK_D2_MRT_ROWS     2   /* MRT.cu:14-22     force moments rows 4/5 in (jy,qx) order        */
#define QK_D3_ZOUHE_RHO    4   /* zouHeInflow.cuh:15-16  f[2] where Zou-He needs f[3]             */
#define QK_D7_IBM_CLIP     8   /* IBM_impl.cuh:20-24,41-44  keep only values > 1e-8               */
#define QK_D8_IBM_2X2      16  /* IBM_impl.cu:19-21,132-134  2x2 stencil under a 4-point delta    */
#define QK_D11_BB_RAW      32  /* bbDomainBoundary.cuh:35-36  raw neighbour test ignores periodic */
#define QK_D9_IBM_ZERO_TARGET 64 /* IBM_impl.cuh:15  target velocity hard-coded 0, IBMBody::velocities dead data */
#define QK_ALL             127

/* BC_flag — lbm_constants.cuh:377-397 */
enum {
    FLUID = 0, BOUNCE_BACK, ZOU_HE_TOP, ZOU_HE_LEFT, ZOU_HE_TOP_LEFT_TOP_INFLOW,
    ZOU_HE_TOP_RIGHT_TOP_INFLOW, CYLINDER, ZG_OUTFLOW, PRESSURE_OUTLET, REGULARIZED_INLET_TOP,
    REGULARIZED_INLET_LEFT, REGULARIZED_BOUNCE_BACK, REGULARIZED_BOUNCE_BACK_CORNER
};

/* lbm_constants.cuh:13-55 */
static const float W[Q] = {4.0f / 9.0f, 1.0f / 9.0f, 1.0f / 9.0f, 1.0f / 9.0f, 1.0f / 9.0f,
                           1.0f / 36.0f, 1.0f / 36.0f, 1.0f / 36.0f, 1.0f / 36.0f};
static const int C[2 * Q] = {0, 0, 1, 0, 0, 1, -1, 0, 0, -1, 1, 1, -1, 1, -1, -1, 1, -1};
static const int OPP[Q] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
static const float Mm[Q * Q] = {
    1, 1, 1, 1, 1, 1, 1, 1, 1,
    -4, -1, -1, -1, -1, 2, 2, 2, 2,
    4, -2, -2, -2, -2, 1, 1, 1, 1,
    0, 1, 0, -1, 0, 1, -1, -1, 1,
    0, -2, 0, 2, 0, 1, -1, -1, 1,
    0, 0, 1, 0, -1, 1, 1, -1, -1,
    0, 0, -2, 0, 2, 1, 1, -1, -1,
    0, 1, -1, 1, -1, 0, 0, 0, 0,
    0, 0, 0, 0, 0, 1, -1, 1, -1};
static const float Mi[Q * Q] = {
    1.0f/9.0f, -1.0f/9.0f,   1.0f/9.0f,   0.0f,       0.0f,        0.0f,       0.0f,        0.0f,      0.0f,
    1.0f/9.0f, -1.0f/36.0f, -1.0f/18.0f,  1.0f/6.0f, -1.0f/6.0f,   0.0f,       0.0f,        1.0f/4.0f, 0.0f,
    1.0f/9.0f, -1.0f/36.0f, -1.0f/18.0f,  0.0f,       0.0f,        1.0f/6.0f, -1.0f/6.0f,  -1.0f/4.0f, 0.0f,
    1.0f/9.0f, -1.0f/36.0f, -1.0f/18.0f, -1.0f/6.0f,  1.0f/6.0f,   0.0f,       0.0f,        1.0f/4.0f, 0.0f,
    1.0f/9.0f, -1.0f/36.0f, -1.0f/18.0f,  0.0f,       0.0f,       -1.0f/6.0f,  1.0f/6.0f,  -1.0f/4.0f, 0.0f,
    1.0f/9.0f,  1.0f/18.0f,  1.0f/36.0f,  1.0f/6.0f,  1.0f/12.0f,  1.0f/6.0f,  1.0f/12.0f,  0.0f,      1.0f/4.0f,
    1.0f/9.0f,  1.0f/18.0f,  1.0f/36.0f, -1.0f/6.0f, -1.0f/12.0f,  1.0f/6.0f,  1.0f/12.0f,  0.0f,     -1.0f/4.0f,
    1.0f/9.0f,  1.0f/18.0f,  1.0f/36.0f, -1.0f/6.0f, -1.0f/12.0f, -1.0f/6.0f, -1.0f/12.0f,  0.0f,      1.0f/4.0f,
    1.0f/9.0f,  1.0f/18.0f,  1.0f/36.0f,  1.0f/6.0f,  1.0f/12.0f, -1.0f/6.0f, -1.0f/12.0f,  0.0f,     -1.0f/4.0f};

typedef struct {
    int nx, ny;
    int periodic_x, periodic_y;       /* streaming.cuh:8-11 macros */
    int coll;                         /* 0 BGK<2>, 1 MRT<2>, 2 CM<2,NoAdapter>, 3 CM<2,OptimalAdapter> */
    float vis, tau, omega;            /* __constant__ vis,tau,omega  lbm.cu:10-12 */
    float S[Q];                       /* __constant__ S  lbm.cu:17 */
    float u_max;                      /* Scenario::u_max handed to the BC functors boundaries.cuh:40,44,70 */
    float force_x, force_y;           /* what Init::apply_forces writes every step (uniform in all 2-D scenarios) */
    int quirks;
    int timestep;
    float *f, *f_back, *f_eq, *rho, *u, *force, *pi_mag, *bc_snapshot;
    int *flags;
    float avg_rho, avg_j, avg_pi;     /* d_moment_avg  lbm.cuh:25-31 */
    /* IBM  IBMManager.cuh:31-52 */
    int np;
    float *pts, *lag_u, *lag_rho, *lag_force, *u_prev, *f_iter;
    float *lag_target;      /* IBMBody::velocities, AoS [i*2+c] (NULL = none given) */
} oracle_t;

/* ------------------------------------------------------------------ lifecycle */
oracle_t *oracle_create(int nx, int ny, int periodic_x, int periodic_y, int coll, float viscosity,
                        const float *S, float u_max, float force_x, float force_y, int quirks) {
    oracle_t *o = (oracle_t *)calloc(1, sizeof(oracle_t));
    size_t n = (size_t)nx * ny;
    o->nx = nx; o->ny = ny; o->periodic_x = periodic_x; o->periodic_y = periodic_y;
    o->coll = coll; o->vis = viscosity;
    o->tau = 3 * viscosity + 0.5f;          /* viscosity_to_tau  lbm_constants.cuh:365-367 */
    o->omega = 1.0f / o->tau;               /* scenario.cuh:37 */
    memcpy(o->S, S, sizeof(float) * Q);
    o->u_max = u_max; o->force_x = force_x; o->force_y = force_y; o->quirks = quirks;
    o->f = (float *)calloc(n * Q, 4); o->f_back = (float *)calloc(n * Q, 4); o->f_eq = (float *)calloc(n * Q, 4);
    o->bc_snapshot = (float *)calloc(n * Q, 4);
    o->rho = (float *)calloc(n, 4); o->u = (float *)calloc(2 * n, 4); o->force = (float *)calloc(2 * n, 4);
    o->pi_mag = (float *)calloc(n, 4); o->flags = (int *)calloc(n, sizeof(int));
    return o;
}

void oracle_destroy(oracle_t *o) {
    if (!o) return;
    free(o->f); free(o->f_back); free(o->f_eq); free(o->rho); free(o->u); free(o->force);
    free(o->pi_mag); free(o->flags); free(o->bc_snapshot);
    free(o->pts); free(o->lag_u); free(o->lag_rho); free(o->lag_force); free(o->u_prev); free(o->f_iter); free(o->lag_target);
    free(o);
}

void oracle_set_flags(oracle_t *o, const int *flags) {      /* setup_boundary_flags  boundaries.cuh:170-197 */
    memcpy(o->flags, flags, sizeof(int) * (size_t)o->nx * o->ny);
}
void oracle_set_omega(oracle_t *o, float tau, float omega) { o->tau = tau; o->omega = omega; }

/* IBMManager::init_and_dispatch / send_to_gpu  IBMManager.cuh:54-109 (points AoS [i*2+c]) */
void oracle_set_markers(oracle_t *o, const float *pts_aos, int np) {
    size_t n = (size_t)o->nx * o->ny;
    free(o->pts); free(o->lag_u); free(o->lag_rho); free(o->lag_force); free(o->u_prev); free(o->f_iter);
    free(o->lag_target); o->lag_target = NULL;
    o->np = np;
    o->pts = (float *)malloc(sizeof(float) * 2 * np);
    memcpy(o->pts, pts_aos, sizeof(float) * 2 * np);
    o->lag_u = (float *)calloc(2 * np, 4); o->lag_rho = (float *)calloc(np, 4); o->lag_force = (float *)calloc(2 * np, 4);
    o->u_prev = (float *)calloc(2 * n, 4); o->f_iter = (float *)calloc(2 * n, 4);
}

/* IBMBody::velocities (IBMBody.cuh:33-45).  The reference uploads them and then forces towards a literal 0
 * (IBM_impl.cuh:15, Appendix A-D9): with QK_D9_IBM_ZERO_TARGET they are ignored, without it u_target = velocities. */
void oracle_set_marker_velocities(oracle_t *o, const float *vel_aos) {
    free(o->lag_target); o->lag_target = NULL;
    if (!vel_aos || o->np == 0) return;
    o->lag_target = (float *)malloc(sizeof(float) * 2 * o->np);
    memcpy(o->lag_target, vel_aos, sizeof(float) * 2 * o->np);
}

/* ------------------------------------------------------------------ equilibrium */
/* LBM<2>::equilibrium_node  equilibrium.cu:5-39 — fp32 inputs, fp64 bracket (A-D19) */
static void equilibrium_node(float *f_eq, float ux, float uy, float rho, size_t node) {
    float u_dot_u = ux * ux + uy * uy;
    float cs = 1.0f / sqrtf(3.0f);
    float cs2 = cs * cs;
    float cs4 = cs2 * cs2;
    for (int q = 0; q < Q; q++) {
        float cu = C[2 * q] * ux + C[2 * q + 1] * uy;
        double cud = (double)cu;
        double br = 1 + 0.5 * (cud * cud) / cs4 - 0.5 * u_dot_u / cs2 + 1.0 * cu / cs2;
        f_eq[node * Q + q] = (float)((double)(W[q] * rho) * br);
    }
}

/* equilibrium_kernel  equilibrium.cuh:20-51 */
static void compute_equilibrium(oracle_t *o) {
    long n = (long)o->nx * o->ny;
#pragma omp parallel for schedule(static)
    for (long node = 0; node < n; node++)
        equilibrium_node(o->f_eq, o->u[2 * node], o->u[2 * node + 1], o->rho[node], (size_t)node);
}

/* init_kernel / init_node  init.cuh:12-43: caller has evaluated the Init functor into rho,u */
void oracle_init(oracle_t *o, const float *rho, const float *u_aos) {
    size_t n = (size_t)o->nx * o->ny;
    memcpy(o->rho, rho, 4 * n);
    memcpy(o->u, u_aos, 8 * n);
    for (size_t i = 0; i < n; i++) { o->force[2 * i] = o->force_x; o->force[2 * i + 1] = o->force_y; }
    compute_equilibrium(o);
    memcpy(o->f, o->f_eq, 4 * n * Q);
    memcpy(o->f_back, o->f_eq, 4 * n * Q);
    o->timestep = 0;
}

/* start from a dumped reference state: f and f_back both given (f_back only matters for A-D1/undelivered slots) */
void oracle_set_populations(oracle_t *o, const float *f, const float *f_back) {
    size_t n = (size_t)o->nx * o->ny;
    memcpy(o->f, f, 4 * n * Q);
    memcpy(o->f_back, f_back, 4 * n * Q);
}

/* ------------------------------------------------------------------ stream */
/* stream_kernel<2> -> LBM<2>::stream_node  streaming.cu:5-33, then swap_buffers lbm.cuh:345-350 */
static void stream_and_swap(oracle_t *o) {
    const int NX = o->nx, NY = o->ny;
    float *f = o->f, *fb = o->f_back;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < NY; y++)
        for (int x = 0; x < NX; x++) {
            size_t node = (size_t)y * NX + x;
            for (int i = 1; i < Q; i++) {              /* starts at 1: A-D1 */
                int xn = x + C[2 * i], yn = y + C[2 * i + 1];
                if (o->periodic_x) xn = (xn + NX) % NX;
                if (o->periodic_y) yn = (yn + NY) % NY;
                if (xn < 0 || xn >= NX || yn < 0 || yn >= NY) continue;
                fb[((size_t)yn * NX + xn) * Q + i] = f[node * Q + i];
            }
            if (!(o->quirks & QK_D1_STALE_F0)) fb[node * Q] = f[node * Q];
        }
    o->f = fb; o->f_back = f;
}

/* ------------------------------------------------------------------ boundary functors */
/* BounceBack<2>::apply  bbDomainBoundary.cuh:22-49 (x=true,y=true  boundaries.cuh:20) */
static void bc_bounce_back(oracle_t *o, float *f, int x, int y) {
    const int NX = o->nx, NY = o->ny;
    for (int i = 1; i < Q; i++) {
        int xn = x + C[2 * i], yn = y + C[2 * i + 1];
        int xb = (xn < 0 || xn >= NX), yb = (yn < 0 || yn >= NY);
        if (!(o->quirks & QK_D11_BB_RAW)) { if (o->periodic_x) xb = 0; if (o->periodic_y) yb = 0; }
        if (xb || yb) f[OPP[i]] = f[i];
    }
}

/* ZouHe::apply_left  zouHeInflow.cuh:9-33 */
static void bc_zou_he_left(oracle_t *o, float *f, float u_lid) {
    const float ux = u_lid;
    float rho;
    if (o->quirks & QK_D3_ZOUHE_RHO)
        rho = (f[0] + f[2] + f[4] + 2 * (f[2] + f[6] + f[7])) / (1.0f - ux);
    else
        rho = (f[0] + f[2] + f[4] + 2 * (f[3] + f[6] + f[7])) / (1.0f - ux);
    f[1] = f[3] + (2.0f / 3.0f) * rho * ux;
    f[5] = f[7] - 0.5f * (f[2] - f[4]) + (1.0f / 6.0f) * rho * ux;
    f[8] = f[6] + 0.5f * (f[2] - f[4]) + (1.0f / 6.0f) * rho * ux;
}

/* ZouHe::apply_top  zouHeInflow.cuh:36-50 */
static void bc_zou_he_top(float *f, float u_lid) {
    const float ux = u_lid, uy = 0.0f;
    float rho = (f[0] + f[1] + f[3] + 2.0f * (f[2] + f[5] + f[6])) / (1.0f + uy);
    f[4] = f[2] - (2.0f / 3.0f) * rho * uy;
    float d13 = f[1] - f[3];
    f[7] = f[5] + 0.5f * d13 - (1.0f / 2.0f) * rho * ux - (1.0f / 6.0f) * rho * uy;
    f[8] = f[6] - 0.5f * d13 + (1.0f / 2.0f) * rho * ux - (1.0f / 6.0f) * rho * uy;
}

/* CylinderBoundary::apply  cylinderBoundary.cuh:22-34 */
static void bc_cylinder(float *f) {
    float saved[Q];
    for (int j = 0; j < Q; j++) saved[j] = f[j];
    for (int i = 0; i < Q; i++) f[OPP[i]] = saved[i];
}

/* ZG_OutflowBoundary<2>::apply  zeroGradientOutflow.cuh:9-59; fn = post-stream f of the interior node */
static void bc_zg_outflow(oracle_t *o, float *f, const float *snap, int x, int y) {
    const int NX = o->nx, NY = o->ny;
    int nrm[2] = {0, 0}, ix = x, iy = y;
    if (x == 0) { nrm[0] = 1; ix = 1; }
    else if (x == NX - 1) { nrm[0] = -1; ix = NX - 2; }
    else if (y == 0) { nrm[1] = 1; iy = 1; }
    else if (y == NY - 1) { nrm[1] = -1; iy = NY - 2; }
    else return;
    const float *fn = snap + ((size_t)iy * NX + ix) * Q;
    for (int i = 0; i < Q; i++) {
        int cdn = C[2 * i] * nrm[0] + C[2 * i + 1] * nrm[1];
        if (cdn > 0) f[i] = fn[i];
    }
}

/* PressureOutlet::apply  pressureOutlet.cuh:7-42 */
static void bc_pressure_outlet(oracle_t *o, float *f, const float *snap, int x, int y) {
    const float *fi = snap + ((size_t)y * o->nx + (x - 1)) * Q;
    float rho_i = 0.0f;
    for (int i = 0; i < Q; i++) rho_i += fi[i];
    float ux = 0.0f, uy = 0.0f;
    for (int i = 0; i < Q; i++) { ux += fi[i] * C[2 * i]; uy += fi[i] * C[2 * i + 1]; }
    ux /= rho_i; uy /= rho_i;
    float target_rho = 1.0f;
    for (int i = 0; i < Q; i++) {
        float cu = C[2 * i] * ux + C[2 * i + 1] * uy;
        float uu = ux * ux + uy * uy;
        float cs2 = 1.0f / 3.0f;
        f[i] = W[i] * target_rho * (1.0f + cu / cs2 + (cu * cu) / (2.0f * cs2 * cs2) - uu / (2.0f * cs2));
    }
}

/* second-order f_eq in pure fp32, as written in regularizedInlet.cuh:24-31 / regularizedBounceBack.cuh:53-58,178-185 */
static float feq32(int q, float rho, float ux, float uy) {
    float cs2 = 1.0f / 3.0f;
    float cu = C[2 * q] * ux + C[2 * q + 1] * uy;
    float uu = ux * ux + uy * uy;
    return W[q] * rho * (1.0f + cu / cs2 + cu * cu / (2.0f * cs2 * cs2) - uu / (2.0f * cs2));
}

/* shared tail of RegularizedInlet::apply_top (:42-67) and RegularizedBounceBack::apply (:68-96) */
static void regularize_from_pi(float *f, const float *feq, float rho, float ux, float uy) {
    float cs2 = 1.0f / 3.0f;
    float Pxx = 0.0f, Pyy = 0.0f, Pxy = 0.0f;
    for (int q = 0; q < Q; q++) {
        float cx = C[2 * q], cy = C[2 * q + 1];
        Pxx += cx * cx * f[q]; Pyy += cy * cy * f[q]; Pxy += cx * cy * f[q];
    }
    Pxx -= cs2 * rho + rho * ux * ux;
    Pyy -= cs2 * rho + rho * uy * uy;
    Pxy -= rho * ux * uy;
    for (int q = 0; q < Q; q++) {
        float cx = C[2 * q], cy = C[2 * q + 1];
        float Qxx = cx * cx - cs2, Qyy = cy * cy - cs2, Qxy = cx * cy;
        float fneq = (W[q] / (2.0f * cs2 * cs2)) * (Qxx * Pxx + Qyy * Pyy + 2.0f * Qxy * Pxy);
        f[q] = feq[q] + fneq;
    }
}

/* RegularizedInlet::apply_top  regularizedInlet.cuh:15-69 */
static void bc_regularized_inlet_top(float *f, float u_lid) {
    const float ux = u_lid, uy = 0.0f;
    float rho = (f[0] + f[1] + f[3] + 2.0f * (f[2] + f[5] + f[6])) / (1.0f + uy);
    float feq[Q];
    for (int q = 0; q < Q; q++) feq[q] = feq32(q, rho, ux, uy);
    for (int q = 0; q < Q; q++)
        if (C[2 * q + 1] < 0) f[q] = feq[q] + (f[OPP[q]] - feq[OPP[q]]);
    regularize_from_pi(f, feq, rho, ux, uy);
}

/* RegularizedBounceBack::apply  regularizedBounceBack.cuh:13-99 */
static void bc_regularized_bb(oracle_t *o, float *f, int x, int y) {
    const int NX = o->nx, NY = o->ny;
    int unk[Q] = {0};
    for (int i = 1; i < Q; i++) {
        if (x == 0) unk[i] = (C[2 * i] > 0);
        else if (x == NX - 1) unk[i] = (C[2 * i] < 0);
        else if (y == 0) unk[i] = (C[2 * i + 1] > 0);
        else if (y == NY - 1) unk[i] = (C[2 * i + 1] < 0);
    }
    float rho = 0.0f;
    for (int i = 0; i < Q; i++)
        if (!unk[i]) { int w = unk[OPP[i]] ? 2 : 1; rho += w * f[i]; }
    const float ux = 0.0f, uy = 0.0f;
    float feq[Q];
    for (int q = 0; q < Q; q++) feq[q] = feq32(q, rho, ux, uy);
    for (int q = 0; q < Q; q++)
        if (unk[q]) f[q] = feq[q] + (f[OPP[q]] - feq[OPP[q]]);
    regularize_from_pi(f, feq, rho, ux, uy);
}

/* RegularizedCornerBounceBack::apply  regularizedBounceBack.cuh:107-221 */
static void bc_regularized_corner(oracle_t *o, float *f, const float *snap, int x, int y) {
    const int NX = o->nx, NY = o->ny;
    int left = (x == 0), right = (x == NX - 1), bottom = (y == 0), top = (y == NY - 1);
    if (!((left || right) && (bottom || top))) return;
    int unk[Q] = {0};
    for (int i = 1; i < Q; i++) {
        int cx = C[2 * i], cy = C[2 * i + 1];
        unk[i] = (left && cx > 0) || (right && cx < 0) || (bottom && cy > 0) || (top && cy < 0);
    }
    int dx = left ? x + 1 : x - 1, dy = bottom ? y + 1 : y - 1;
    dx = dx < NX - 2 ? dx : NX - 2; dx = dx > 1 ? dx : 1;       /* max(1, min(d, N-2)) :140-141 */
    dy = dy < NY - 2 ? dy : NY - 2; dy = dy > 1 ? dy : 1;
    const float *fd = snap + ((size_t)dy * NX + dx) * Q;
    float rho = 0.0f;
    for (int i = 0; i < Q; i++) rho += fd[i];
    const float ux = 0.0f, uy = 0.0f;
    for (int i = 0; i < Q; i++)
        if (unk[i]) f[i] = feq32(i, rho, ux, uy) - (f[OPP[i]] - feq32(OPP[i], rho, ux, uy));   /* minus sign :164 */
    /* regularize_distributions :192-221 — Pi from f - f_eq */
    float cs2 = 1.0f / 3.0f;
    float Pxx = 0.0f, Pyy = 0.0f, Pxy = 0.0f;
    for (int q = 0; q < Q; q++) {
        float fneq = f[q] - feq32(q, rho, ux, uy);
        float cx = C[2 * q], cy = C[2 * q + 1];
        Pxx += cx * cx * fneq; Pyy += cy * cy * fneq; Pxy += cx * cy * fneq;
    }
    for (int q = 0; q < Q; q++) {
        float cx = C[2 * q], cy = C[2 * q + 1];
        float Qxx = cx * cx - cs2, Qyy = cy * cy - cs2, Qxy = cx * cy;
        float fneq = (W[q] / (2.0f * cs2 * cs2)) * (Qxx * Pxx + Qyy * Pyy + 2.0f * Qxy * Pxy);
        f[q] = feq32(q, rho, ux, uy) + fneq;
    }
}

/* boundaries_kernel_2D  boundaries.cuh:10-85.  Neighbour reads (ZG / PRESSURE / corner) go to a
 * snapshot of the post-stream field: in the reference they read FLUID interior nodes, which no
 * thread of that kernel modifies, so the snapshot is what it sees. */
static void apply_boundaries(oracle_t *o) {
    const int NX = o->nx, NY = o->ny;
    size_t n = (size_t)NX * NY;
    int need_snap = 0;
    for (size_t i = 0; i < n && !need_snap; i++) {
        int fl = o->flags[i];
        need_snap = (fl == ZG_OUTFLOW || fl == PRESSURE_OUTLET || fl == REGULARIZED_BOUNCE_BACK_CORNER);
    }
    if (need_snap) memcpy(o->bc_snapshot, o->f, 4 * n * Q);
    const float *snap = o->bc_snapshot;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < NY; y++)
        for (int x = 0; x < NX; x++) {
            size_t node = (size_t)y * NX + x;
            float *f = o->f + node * Q;
            switch (o->flags[node]) {
            case FLUID: break;
            case BOUNCE_BACK: bc_bounce_back(o, f, x, y); break;
            case ZOU_HE_TOP: bc_zou_he_top(f, o->u_max); break;
            case ZOU_HE_LEFT: bc_zou_he_left(o, f, o->u_max); break;
            case CYLINDER: bc_cylinder(f); break;
            case ZG_OUTFLOW: bc_zg_outflow(o, f, snap, x, y); break;
            case PRESSURE_OUTLET: bc_pressure_outlet(o, f, snap, x, y); break;
            case REGULARIZED_INLET_TOP: bc_regularized_inlet_top(f, o->u_max); break;
            case REGULARIZED_BOUNCE_BACK: bc_regularized_bb(o, f, x, y); break;
            case REGULARIZED_BOUNCE_BACK_CORNER: bc_regularized_corner(o, f, snap, x, y); break;
            default: break;     /* "Unknown Flag" printf, no-op  boundaries.cuh:78-81 */
            }
        }
}

/* ------------------------------------------------------------------ macroscopics */
/* uncorrected_macroscopics_kernel<2>  macroscopics.cu:5-38 */
static void uncorrected_macroscopics(oracle_t *o) {
    long n = (long)o->nx * o->ny;
#pragma omp parallel for schedule(static)
    for (long node = 0; node < n; node++) {
        const float *f = o->f + node * Q;
        float rho = 0.0f, ux = 0.0f, uy = 0.0f, pi[3] = {0.0f, 0.0f, 0.0f};
        for (int i = 0; i < Q; i++) {
            float fi = f[i];
            rho += fi;
            ux += fi * C[2 * i];
            uy += fi * C[2 * i + 1];
            pi[0] += fi * C[2 * i] * C[2 * i];
            pi[1] += fi * C[2 * i] * C[2 * i + 1];
            pi[2] += fi * C[2 * i + 1] * C[2 * i + 1];
        }
        ux *= 1.0f / rho;
        uy *= 1.0f / rho;
        o->rho[node] = rho; o->u[2 * node] = ux; o->u[2 * node + 1] = uy;
        o->pi_mag[node] = sqrtf(pi[0] * pi[0] + 2.0f * pi[1] * pi[1] + pi[2] * pi[2]);
    }
}

/* reset_forces_kernel -> Init::apply_forces  macroscopics.cuh:13-48 */
static void reset_forces(oracle_t *o) {
    long n = (long)o->nx * o->ny;
#pragma omp parallel for schedule(static)
    for (long node = 0; node < n; node++) { o->force[2 * node] = o->force_x; o->force[2 * node + 1] = o->force_y; }
}

/* correct_macroscopics_kernel<2>  macroscopics.cu:99-110 */
static void correct_macroscopics(oracle_t *o) {
    long n = (long)o->nx * o->ny;
#pragma omp parallel for schedule(static)
    for (long node = 0; node < n; node++) {
        o->u[2 * node] += 0.5f * o->force[2 * node] / o->rho[node];
        o->u[2 * node + 1] += 0.5f * o->force[2 * node + 1] / o->rho[node];
    }
}

/* update_avg_mag<2>  macroscopics.cuh:51-120 + host part :161-177.  16x16 blocks, smem tree, then one
 * atomicAdd of block_sum/domain_size per block; block order (atomics) taken row-major here (A-D17). */
static void update_avg_mag(oracle_t *o) {
    const int NX = o->nx, NY = o->ny, B = 16;
    int gx = (NX + B - 1) / B, gy = (NY + B - 1) / B;
    float domain = (float)(NX * NY);
    float *part = (float *)malloc(sizeof(float) * 3 * (size_t)gx * gy);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < gx * gy; b++) {
        int bx = b % gx, by = b / gx;
        float s0[256], s1[256], s2[256];
        for (int ty = 0; ty < B; ty++)
            for (int tx = 0; tx < B; tx++) {
                int x = bx * B + tx, y = by * B + ty, tid = ty * B + tx;
                float lr = 0.0f, lj = 0.0f, lp = 0.0f;
                if (x < NX && y < NY) {
                    size_t node = (size_t)y * NX + x;
                    float ux = o->u[2 * node], uy = o->u[2 * node + 1];
                    lr = o->rho[node];
                    lj = lr * sqrtf(ux * ux + uy * uy);
                    lp = o->pi_mag[node];
                }
                s0[tid] = lr; s1[tid] = lj; s2[tid] = lp;
            }
        for (int s = 128; s > 0; s >>= 1)
            for (int tid = 0; tid < s; tid++) { s0[tid] += s0[tid + s]; s1[tid] += s1[tid + s]; s2[tid] += s2[tid + s]; }
        part[3 * b] = s0[0] / domain; part[3 * b + 1] = s1[0] / domain; part[3 * b + 2] = s2[0] / domain;
    }
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
    for (int b = 0; b < gx * gy; b++) { a0 += part[3 * b]; a1 += part[3 * b + 1]; a2 += part[3 * b + 2]; }
    free(part);
    o->avg_rho = a0; o->avg_j = a1; o->avg_pi = a2;
}

/* ------------------------------------------------------------------ IBM */
/* delta4 / kernel2D  IBMUtils.cuh:23-43 */
static float delta4(float r) {
    float rabs = fabsf(r);
    if (rabs < 1.0f) return 0.125f * (3.0f - 2.0f * rabs + sqrtf(1 + 4.0f * rabs - 4.0f * r * r));
    else if (rabs < 2.0f) return 0.125f * (5.0f - 2.0f * rabs - sqrtf(-7.0f + 12.0f * rabs - 4.0f * r * r));
    else return 0.0f;
}
static float kernel2D(float dx, float dy) { return delta4(dx) * delta4(dy); }

/* IBMManager<2>::multi_direct  IBMManager.cuh:222-252 with the five kernels it launches */
static void ibm_multi_direct(oracle_t *o) {
    if (o->np == 0) return;
    const int NX = o->nx, NY = o->ny, np = o->np;
    size_t n = (size_t)NX * NY;
    const int lo = (o->quirks & QK_D8_IBM_2X2) ? 0 : -1, hi = (o->quirks & QK_D8_IBM_2X2) ? 2 : 3;
    memcpy(o->u_prev, o->u, 8 * n);                               /* :227 */
    for (int iter = 0; iter < 3; iter++) {                        /* ITER_MAX 3  :8 */
        memset(o->f_iter, 0, 8 * n);                              /* :235 */
        /* interpolate_velocities_kernel<2>  IBM_impl.cu:7-51 (lag arrays SoA [c*np+i] IBMUtils.cuh:67-71) */
        for (int k = 0; k < np; k++) {
            float px = o->pts[2 * k], py = o->pts[2 * k + 1];
            float gx = floorf(px), gy = floorf(py);
            float ux = 0.0f, uy = 0.0f, rho = 0.0f;
            for (int i = lo; i < hi; i++)
                for (int j = lo; j < hi; j++) {
                    int nx_ = (int)(gx + i), ny_ = (int)(gy + j);
                    if (nx_ >= NX || nx_ < 0 || ny_ >= NY || ny_ < 0) continue;
                    float dx = px - nx_, dy = py - ny_;
                    float kk = kernel2D(dx, dy);
                    size_t ui = (size_t)nx_ + (size_t)ny_ * NX;
                    rho += kk * o->rho[ui];
                    ux += kk * o->u_prev[2 * ui];
                    uy += kk * o->u_prev[2 * ui + 1];
                }
            o->lag_u[k] = ux; o->lag_u[np + k] = uy; o->lag_rho[k] = rho;
        }
        /* compute_lagrangian_kernel  IBM_impl.cuh:9-26 */
        for (int k = 0; k < np; k++)
            for (int d = 0; d < 2; d++) {
                float uu = o->lag_u[d * np + k];
                float ut = (o->lag_target && !(o->quirks & QK_D9_IBM_ZERO_TARGET)) ? o->lag_target[2 * k + d] : 0.0f;   /* :15 u_target */
                float force = 2.0f * o->lag_rho[k] * (ut - uu);
                if (o->quirks & QK_D7_IBM_CLIP) force = force > 1e-8f ? force : 0.0f;
                o->lag_force[d * np + k] = force;
            }
        /* spread_forces_kernel<2>  IBM_impl.cu:122-154 (atomicAdd order: marker index, A-D17) */
        for (int k = 0; k < np; k++) {
            float px = o->pts[2 * k], py = o->pts[2 * k + 1];
            float gx = floorf(px), gy = floorf(py);
            for (int i = lo; i < hi; i++)
                for (int j = lo; j < hi; j++) {
                    int nx_ = (int)(gx + i), ny_ = (int)(gy + j);
                    if (nx_ >= NX || nx_ < 0 || ny_ >= NY || ny_ < 0) continue;
                    float dx = px - (gx + i), dy = py - (gy + j);
                    float kk = kernel2D(dx, dy);
                    float fx = kk * o->lag_force[k], fy = kk * o->lag_force[np + k];
                    size_t ni = (size_t)ny_ * NX + nx_;
                    o->f_iter[2 * ni] += fx; o->f_iter[2 * ni + 1] += fy;
                }
        }
        /* correct_velocities_kernel  IBM_impl.cuh:30-46 and accumulate_forces_kernel :50-68 */
#pragma omp parallel for schedule(static)
        for (long node = 0; node < (long)n; node++)
            for (int d = 0; d < 2; d++) {
                float cu = o->u_prev[2 * node + d] + o->f_iter[2 * node + d] / (2.0f * o->rho[node]);
                if (o->quirks & QK_D7_IBM_CLIP) cu = ((double)cu > 1e-8) ? cu : 0.0f;
                o->u_prev[2 * node + d] = cu;
                o->force[2 * node + d] += o->f_iter[2 * node + d];
            }
    }
}

/* ------------------------------------------------------------------ collision */
/* BGK<2>::apply  BGK/BGK.cuh:13-51 */
static void collide_bgk(oracle_t *o, size_t node) {
    float *f = o->f + node * Q; const float *feq = o->f_eq + node * Q;
    const float omega = o->omega;
    for (int q = 0; q < Q; q++) {
        float cx = C[2 * q], cy = C[2 * q + 1];
        float fx = o->force[2 * node], fy = o->force[2 * node + 1];
        float ux = o->u[2 * node], uy = o->u[2 * node + 1];
        float cs2 = 1.0f / 3.0f;
        float force_term = W[q] * ((1.0f - 0.5f * omega) * ((cx - ux) / cs2 + (cx * ux + cy * uy) * cx / (cs2 * cs2)) * fx +
                                   (1.0f - 0.5f * omega) * ((cy - uy) / cs2 + (cx * ux + cy * uy) * cy / (cs2 * cs2)) * fy);
        f[q] = f[q] - omega * (f[q] - feq[q]) + force_term;
    }
}

/* MRT<2>::compute_forcing_term + MRT<2>::apply  MRT/MRT.cu:4-76 */
static void collide_mrt(oracle_t *o, size_t node) {
    float *f = o->f + node * Q; const float *feq = o->f_eq + node * Q;
    float fx = o->force[2 * node], fy = o->force[2 * node + 1];
    float ux = o->u[2 * node], uy = o->u[2 * node + 1];
    float F[Q], m[Q], meq[Q], mp[Q];
    F[0] = 0.0f;
    F[1] = 6.0f * (fx * ux + fy * uy);
    F[2] = -6.0f * (fx * ux + fy * uy);
    F[3] = fx;
    if (o->quirks & QK_D2_MRT_ROWS) { F[4] = fy; F[5] = -fx; }     /* rows in (jy,qx) order against M's (qx,jy) */
    else { F[4] = -fx; F[5] = fy; }
    F[6] = -fy;
    F[7] = 2.0f * (fx * ux - fy * uy);
    F[8] = fx * uy + fy * ux;
    for (int k = 0; k < Q; k++) {
        m[k] = 0.0f; meq[k] = 0.0f;
        for (int i = 0; i < Q; i++) { m[k] += Mm[k * Q + i] * f[i]; meq[k] += Mm[k * Q + i] * feq[i]; }
        float src = (1.0f - o->S[k] / 2.0f) * F[k];
        mp[k] = m[k] - o->S[k] * (m[k] - meq[k]) + src;
    }
    for (int k = 0; k < Q; k++) {
        float acc = 0.0f;
        for (int i = 0; i < Q; i++) acc += Mi[k * Q + i] * mp[i];
        f[k] = acc;
    }
}

/* CM<2,Adapter>::cm_matrix_inverse  CM/CM.cuh:141-250 (generated by codegen/cm_matrix_inv.py) */
static void cm_matrix_inverse(float *T, float ux, float uy) {
    float ux2 = ux * ux, uy2 = uy * uy, uxuy = ux * uy;
    float ux2uy = uy * (ux * ux), uxuy2 = ux * (uy * uy), ux2uy2 = (ux * ux) * (uy * uy);
    float x3 = -uxuy2, x5 = -uy;
    T[0] = -ux2 + ux2uy2 - uy2 + 1.0f; T[1] = -2.0f * ux + 2.0f * uxuy2; T[2] = 2.0f * ux2uy + 2.0f * x5;
    T[3] = 0.5f * ux2 + 0.5f * uy2 - 1.0f; T[4] = -0.5f * ux2 + 0.5f * uy2; T[5] = 4.0f * uxuy;
    T[6] = 2.0f * uy; T[7] = 2.0f * ux; T[8] = 1.0f;

    T[9] = 0.5f * ux + 0.5f * ux2 - 0.5f * ux2uy2 + 0.5f * x3; T[10] = ux - 0.5f * uy2 + x3 + 0.5f; T[11] = -ux2uy - uxuy;
    T[12] = -0.25f * ux - 0.25f * ux2 - 0.25f * uy2 + 0.25f; T[13] = 0.25f * ux + 0.25f * ux2 - 0.25f * uy2 + 0.25f;
    T[14] = -2.0f * uxuy + x5; T[15] = x5; T[16] = -ux - 0.5f; T[17] = -0.5f;

    T[18] = -0.5f * ux2uy - 0.5f * ux2uy2 + 0.5f * uy + 0.5f * uy2; T[19] = -uxuy + x3; T[20] = -0.5f * ux2 - ux2uy + uy + 0.5f;
    T[21] = -0.25f * ux2 - 0.25f * uy2 + 0.25f * x5 + 0.25f; T[22] = 0.25f * ux2 - 0.25f * uy2 + 0.25f * x5 - 0.25f;
    T[23] = -ux - 2.0f * uxuy; T[24] = x5 - 0.5f; T[25] = -ux; T[26] = -0.5f;

    T[27] = -0.5f * ux + 0.5f * ux2 - 0.5f * ux2uy2 + 0.5f * uxuy2; T[28] = ux + 0.5f * uy2 + x3 - 0.5f; T[29] = -ux2uy + uxuy;
    T[30] = 0.25f * ux - 0.25f * ux2 - 0.25f * uy2 + 0.25f; T[31] = -0.25f * ux + 0.25f * ux2 - 0.25f * uy2 + 0.25f;
    T[32] = -2.0f * uxuy + uy; T[33] = x5; T[34] = 0.5f - ux; T[35] = -0.5f;

    T[36] = 0.5f * ux2uy - 0.5f * ux2uy2 + 0.5f * uy2 + 0.5f * x5; T[37] = uxuy + x3; T[38] = 0.5f * ux2 - ux2uy + uy - 0.5f;
    T[39] = -0.25f * ux2 + 0.25f * uy - 0.25f * uy2 + 0.25f; T[40] = 0.25f * ux2 + 0.25f * uy - 0.25f * uy2 - 0.25f;
    T[41] = ux - 2.0f * uxuy; T[42] = x5 + 0.5f; T[43] = -ux; T[44] = -0.5f;

    T[45] = 0.25f * ux2uy + 0.25f * ux2uy2 + 0.25f * uxuy + 0.25f * uxuy2; T[46] = 0.5f * uxuy + 0.5f * uxuy2 + 0.25f * uy + 0.25f * uy2;
    T[47] = 0.25f * ux + 0.25f * ux2 + 0.5f * ux2uy + 0.5f * uxuy; T[48] = 0.125f * ux + 0.125f * ux2 + 0.125f * uy + 0.125f * uy2;
    T[49] = -0.125f * ux - 0.125f * ux2 + 0.125f * uy + 0.125f * uy2; T[50] = 0.5f * ux + uxuy + 0.5f * uy + 0.25f;
    T[51] = 0.5f * uy + 0.25f; T[52] = 0.5f * ux + 0.25f; T[53] = 0.25f;

    T[54] = 0.25f * ux2uy + 0.25f * ux2uy2 - 0.25f * uxuy + 0.25f * x3; T[55] = 0.5f * uxuy + 0.5f * uxuy2 - 0.25f * uy2 + 0.25f * x5;
    T[56] = -0.25f * ux + 0.25f * ux2 + 0.5f * ux2uy - 0.5f * uxuy; T[57] = -0.125f * ux + 0.125f * ux2 + 0.125f * uy + 0.125f * uy2;
    T[58] = 0.125f * ux - 0.125f * ux2 + 0.125f * uy + 0.125f * uy2; T[59] = 0.5f * ux + uxuy + 0.5f * x5 - 0.25f;
    T[60] = 0.5f * uy + 0.25f; T[61] = 0.5f * ux - 0.25f; T[62] = 0.25f;

    T[63] = -0.25f * ux2uy + 0.25f * ux2uy2 + 0.25f * uxuy + 0.25f * x3; T[64] = -0.5f * uxuy + 0.5f * uxuy2 + 0.25f * uy - 0.25f * uy2;
    T[65] = 0.25f * ux - 0.25f * ux2 + 0.5f * ux2uy - 0.5f * uxuy; T[66] = -0.125f * ux + 0.125f * ux2 + 0.125f * uy2 + 0.125f * x5;
    T[67] = 0.125f * ux - 0.125f * ux2 + 0.125f * uy2 + 0.125f * x5; T[68] = -0.5f * ux + uxuy + 0.5f * x5 + 0.25f;
    T[69] = 0.5f * uy - 0.25f; T[70] = 0.5f * ux - 0.25f; T[71] = 0.25f;

    T[72] = -0.25f * ux2uy + 0.25f * ux2uy2 - 0.25f * uxuy + 0.25f * uxuy2; T[73] = -0.5f * uxuy + 0.5f * uxuy2 + 0.25f * uy2 + 0.25f * x5;
    T[74] = -0.25f * ux - 0.25f * ux2 + 0.5f * ux2uy + 0.5f * uxuy; T[75] = 0.125f * ux + 0.125f * ux2 + 0.125f * uy2 + 0.125f * x5;
    T[76] = -0.125f * ux - 0.125f * ux2 + 0.125f * uy2 + 0.125f * x5; T[77] = -0.5f * ux + uxuy + 0.5f * uy - 0.25f;
    T[78] = 0.5f * uy - 0.25f; T[79] = 0.5f * ux + 0.25f; T[80] = 0.25f;
}

/* OptimalAdapter::compute_higher_order_relaxation  adapters.cuh:48-111 */
static float optimal_adapter_rate(float rho, float j_mag, float pi_mag, float a_rho, float a_j, float a_pi) {
    float sp[4] = {rho / a_rho, j_mag / a_j, pi_mag / a_pi, 1.0f};
    float theta[4] = {0.0003f, -0.00775f, 0.00016f, 0.0087f};
    float tau_star = 0.0f;
    for (int i = 0; i < 4; i++) tau_star += theta[i] * sp[i];
    tau_star = tau_star > 0.0f ? tau_star : 0.005f;
    tau_star = fminf(tau_star, 1.5f);
    return 1.0f / (3.0f * tau_star + 0.5f);
}

/* CM<2,Adapter>::apply  CM/CM.cuh:27-139 */
static void collide_cm(oracle_t *o, size_t node, int optimal) {
    float *f = o->f + node * Q;
    float ux = o->u[2 * node], uy = o->u[2 * node + 1];
    float Fx = o->force[2 * node], Fy = o->force[2 * node + 1];
    float rho = 0.0f;
    for (int i = 0; i < Q; i++) rho += f[i];
    float k[Q] = {0.0f}, keq[Q], kp[Q], pi[3] = {0.0f, 0.0f, 0.0f}, T[Q * Q];
    k[0] = rho;
    for (int i = 0; i < Q; i++) {
        float ccx = C[2 * i] - ux, ccy = C[2 * i + 1] - uy;
        float ccx2 = ccx * ccx, ccy2 = ccy * ccy;
        float fi = f[i];
        k[1] += fi * (ccx);
        k[2] += fi * (ccy);
        k[3] += fi * (ccx2 + ccy2);
        k[4] += fi * (ccx2 - ccy2);
        k[5] += fi * ccx * ccy;
        k[6] += fi * ccx2 * ccy;
        k[7] += fi * ccx * ccy2;
        k[8] += fi * ccx2 * ccy2;
        pi[0] += fi * C[2 * i] * C[2 * i];
        pi[1] += fi * C[2 * i] * C[2 * i + 1];
        pi[2] += fi * C[2 * i + 1] * C[2 * i + 1];
    }
    float pi_mag = sqrtf(pi[0] * pi[0] + 2.0f * pi[1] * pi[1] + pi[2] * pi[2]);
    const float cs2 = 1.0f / 3.0f;
    keq[0] = rho; keq[1] = 0.0f; keq[2] = 0.0f; keq[3] = 2.0f * rho * cs2; keq[4] = 0.0f; keq[5] = 0.0f;
    keq[6] = 0.0f; keq[7] = 0.0f; keq[8] = rho * cs2 * cs2;
    float F[Q] = {0.0f, Fx, Fy, 0.0f, 0.0f, 0.0f, Fy * cs2, Fx * cs2, 0.0f};
    float j_mag = sqrtf(ux * ux + uy * uy) * rho;
    float hi_rate = 0.0f;
    if (optimal) hi_rate = optimal_adapter_rate(rho, j_mag, pi_mag, o->avg_rho, o->avg_j, o->avg_pi);
    for (int i = 0; i < Q; i++) {
        float rate = (optimal && i > 5) ? hi_rate : o->S[i];       /* AdapterBase::is_higher_order adapters.cuh:8-13 */
        kp[i] = k[i] - rate * (k[i] - keq[i]) + (1.0f - 0.5f * rate) * F[i];
    }
    cm_matrix_inverse(T, ux, uy);
    for (int i = 0; i < Q; i++) {
        float acc = 0.0f;
        for (int j = 0; j < Q; j++) acc += T[i * Q + j] * kp[j];
        f[i] = acc;
    }
}

/* collide_kernel<Op,2>  collision.cuh:16-65 */
static void collide(oracle_t *o) {
    long n = (long)o->nx * o->ny;
#pragma omp parallel for schedule(static)
    for (long node = 0; node < n; node++) {
        switch (o->coll) {
        case 0: collide_bgk(o, (size_t)node); break;
        case 1: collide_mrt(o, (size_t)node); break;
        case 2: collide_cm(o, (size_t)node, 0); break;
        default: collide_cm(o, (size_t)node, 1); break;
        }
    }
}

/* ------------------------------------------------------------------ time step: main.cu:96-114 */
void oracle_step(oracle_t *o, int nsteps) {
    for (int s = 0; s < nsteps; s++) {
        o->timestep++;                     /* increase_ts  lbm.cuh:142-146 */
        stream_and_swap(o);                /* main.cu:98-99 */
        apply_boundaries(o);               /* :101 */
        uncorrected_macroscopics(o);       /* :103 */
        reset_forces(o);                   /* :106 */
        ibm_multi_direct(o);               /* :107 */
        correct_macroscopics(o);           /* :110 (kernel) */
        update_avg_mag(o);                 /* :110 (second half of LBM::correct_macroscopics) */
        compute_equilibrium(o);            /* :113 */
        collide(o);                        /* :114 */
    }
}

/* ------------------------------------------------------------------ accessors */
void oracle_get_macroscopics(const oracle_t *o, float *rho, float *u_aos) {    /* update_macroscopics lbm.cuh:148-154 */
    size_t n = (size_t)o->nx * o->ny;
    memcpy(rho, o->rho, 4 * n); memcpy(u_aos, o->u, 8 * n);
}
void oracle_get_populations(const oracle_t *o, float *f_aos) { memcpy(f_aos, o->f, 4 * (size_t)o->nx * o->ny * Q); }
void oracle_get_populations_back(const oracle_t *o, float *f_aos) { memcpy(f_aos, o->f_back, 4 * (size_t)o->nx * o->ny * Q); }
void oracle_get_force(const oracle_t *o, float *force_aos) { memcpy(force_aos, o->force, 8 * (size_t)o->nx * o->ny); }
void oracle_get_moment_avg(const oracle_t *o, float *out3) { out3[0] = o->avg_rho; out3[1] = o->avg_j; out3[2] = o->avg_pi; }
int oracle_timestep(const oracle_t *o) { return o->timestep; }
double oracle_total_mass(const oracle_t *o) {
    double s = 0.0; size_t n = (size_t)o->nx * o->ny * Q;
    for (size_t i = 0; i < n; i++) s += o->f[i];
    return s;
}
int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers: the timed CPU legs of bench.py ask for the host's cores explicitly */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------ scenario functors (host side) */
/* TaylorGreenInit::operator()  taylorGreenFunctors.cuh:25-47  (u_max already divided by SCALE, :11-13) */
void oracle_init_taylor_green(int nx, int ny, float nu, float u_max, float *rho, float *u_aos) {
    for (int node = 0; node < nx * ny; node++) {
        const float x = (node % nx) + 0.5f, y = (node / nx) + 0.5f;
        const float rho0 = 1.0f;
        const float kx = (float)(2.0f * M_PI / nx), ky = (float)(2.0f * M_PI / ny);
        const float td = 1.0f / (nu * (kx * kx + ky * ky));
        const float t = 0.0f;
        float ux = -u_max * sqrtf(ky / kx) * cosf(kx * x) * sinf(ky * y) * expf(-t / td);
        float uy = u_max * sqrtf(kx / ky) * sinf(kx * x) * cosf(ky * y) * expf(-t / td);
        float P = -0.25f * rho0 * u_max * u_max * ((ky / kx) * cosf(2 * kx * x) + (kx / ky) * cosf(2 * ky * y)) * expf(-2 * t / td);
        rho[node] = rho0 + 3.0f * P;
        u_aos[2 * node] = ux; u_aos[2 * node + 1] = uy;
    }
}

/* TaylorGreenValidation::operator()  taylorGreenFunctors.cuh:66-81 (u_max = 0.04f/SCALE hard-coded :72) */
void oracle_taylor_green_analytic(int nx, int ny, float nu, float u_max, float t, float *u_aos) {
    for (int yn = 0; yn < ny; yn++)
        for (int xn = 0; xn < nx; xn++) {
            const float x = xn + 0.5f, y = yn + 0.5f;
            const float kx = (float)(2.0f * M_PI / nx), ky = (float)(2.0f * M_PI / ny);
            const float td = 1.0f / (nu * (kx * kx + ky * ky));
            const float decay = expf(-t / td);
            u_aos[2 * (yn * nx + xn)] = -u_max * sqrtf(ky / kx) * cosf(kx * x) * sinf(ky * y) * decay;
            u_aos[2 * (yn * nx + xn) + 1] = u_max * sqrtf(kx / ky) * sinf(kx * x) * cosf(ky * y) * decay;
        }
}

/* PoiseuilleInit::apply_forces  poiseuilleFunctors.cuh:37 */
float oracle_poiseuille_force(float vis, float u_max, int ny) { return 8.0f * vis * u_max / (ny * ny); }

/* create_cylinder  IBM_generators.cu:5-25 */
void oracle_create_cylinder(float cx, float cy, float r, int num_pts, float *pts_aos) {
    float angle = (float)(2 * M_PI / num_pts);
    for (int i = 0; i < num_pts; i++) {
        pts_aos[2 * i] = cx + r * cosf(i * angle);
        pts_aos[2 * i + 1] = cy + r * sinf(i * angle);
    }
}
