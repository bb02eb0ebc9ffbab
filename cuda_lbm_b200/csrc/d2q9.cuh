// d2q9.cuh — device-side building blocks of the fused D2Q9 step (sm_100a).
//
// What one reference time step computes for a cell is specified in SURVEY.md Appendix D
// (restating src/main.cu:96-114 of Carabalone/cuda-lbm).  The reference runs it as nine
// full-grid kernels over three population buffers; here it is one register-resident
// function chain per cell:  pull -> boundary -> moments -> force -> collide -> store,
// over ONE in-place SoA population buffer addressed with the AA pattern:
//
//   even step (local):      g_q(x) = A[q][x]               store f*_q -> A[opp q][x]
//   odd  step (neighbour):  g_q(x) = A[opp q][x - c_q]     store f*_q -> A[q][x + c_q]
//
// Every cell reads exactly the nine slots it later overwrites, so no second buffer and no
// inter-thread hazard exists.  Arithmetic follows the reference's formulas (cited per
// function) in pure fp32 with FMA contraction; parity is to fp32 round-off (tests state it).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "collide.cuh"

namespace lbm {

// BC_flag values — reference src/core/lbm_constants.cuh:377-397
enum : int {
    FLUID = 0, BOUNCE_BACK = 1, ZOU_HE_TOP = 2, ZOU_HE_LEFT = 3, CYLINDER = 6, ZG_OUTFLOW = 7,
    PRESSURE_OUTLET = 8, REGULARIZED_INLET_TOP = 9, REGULARIZED_BOUNCE_BACK = 11,
    REGULARIZED_BOUNCE_BACK_CORNER = 12
};
constexpr uint8_t FLAG_IBM = 0x80;   // node lies in a marker stencil: force comes from ibm_force[]
constexpr uint8_t FLAG_OWNED = 0x40; // FLUID node handed to the general kernel: a boundary node reads its post-stream populations (nbr_gather)
constexpr uint8_t FLAG_MASK = 0x1f;

constexpr int COLCLASS_MAX = 256;     // segment columns covered by Params::colclass (nx <= 32768)

struct Params {
    float* A[Q];            // slot planes, each (ny_local+2) rows of nx floats; row 0 / ny_local+1 are ghost rows
    float* A0[2];           // rest-population plane used at odd / even timesteps (same pointer unless QK_D1)
    int nx, ny;             // global grid
    int y0, nyl;            // slab rows [y0, y0+nyl)
    int px, py;             // periodic axes
    int wrap_y;             // world==1 && periodic_y: y wraps inside the slab (ghost rows unused)
    int t;                  // timestep being computed (1-based); parity selects the AA phase
    int quirks;
    int coll;
    const uint8_t* flags;   // nyl*nx, nullptr = all FLUID and no IBM
    float omega;
    float S[Q];
    float u_max;
    float fx, fy;           // uniform body force
    const float2* force_plane;   // optional per-node body force (local nodes)
    float* ring;            // [2][perim][9] post-collision populations of domain-edge nodes of the last two steps
    int perim;
    const long long* nbr_nodes; const float* nbr_g; int nbr_count;   // neighbour-reading BC nodes (sorted global ids) and the neighbour's post-stream g
    const long long* ibm_nodes; const float2* ibm_force; int ibm_count;
    const float* avg;       // 3 floats: grid means of rho, rho|u|, |Pi| (OptimalAdapter)
    float* partials;        // per-block partial sums (3 per block) or nullptr
    float* rho_out; float2* u_out;     // macroscopic output of this step (nullptr = none)
    // segments: 128 consecutive cells of one row.  segmask[yl*nsx + sx]: 0 = every cell is plain stream + collide (vectorised
    // kernels), 1 = every cell takes the general path (boundary flags, IBM nodes, per-node force, non-periodic domain edge), 2 = mixed:
    // the vectorised kernels skip exactly the general cells.  The general (scalar) kernel is launched over the segments listed in
    // gen_list (type 1) and over the cells listed in gen_cells (the general cells of type-2 segments, local node ids).
    const uint8_t* segmask; int nsx; const int* gen_list;
    // rows [pure_y0, pure_y1) x segments [pure_s0, pure_s1): a rectangle in which EVERY segment is class 0 (the largest one, found on the
    // host).  A warp inside it does not read segmask at all: waiting for that byte before it requests its populations made every warp
    // live one L2 round trip longer (the cavity's vector kernels ran at 219 / 249 us against 197 / 213 us for the same operator on
    // the periodic box), and ptxas moves the test in front of the loads however the source orders them.
    int pure_y0, pure_y1, pure_s0, pure_s1;
    // and for the rows of that rectangle, the class of every segment COLUMN whose class does not change over those rows (255: it does,
    // look it up) — the cavity's two wall columns are mixed segments in every interior row, and a block is held by its slowest warp
    uint8_t colclass[COLCLASS_MAX];
    const long long* gen_cells; long long gen_cell_count;
    // direct y-slab coupling over NVLink peer memory: [0] = the lower neighbour's top edge row, [1] = the upper neighbour's
    // bottom edge row, addressed as peer[s] + plane * peer_plane[s] + peer_off[s] + x.  nullptr = use this slab's ghost rows.
    float* peer[2]; long long peer_plane[2]; long long peer_off[2];
    long long plane;        // floats per slot plane of this slab: A[q] == A[0] + q * plane
};
constexpr int SEG = 128;
constexpr uint8_t SEG_LOOKUP = 255;

__device__ __forceinline__ long long rowoff(const Params& p, int yl) { return (long long)(yl + 1) * p.nx; }
// start of row yl of a slot plane; rows -1 and nyl resolve to the neighbour slab's edge row when it is peer-mapped
__device__ __forceinline__ float* row_ptr(const Params& p, int plane, int yl) {
    if (yl < 0 && p.peer[0]) return p.peer[0] + plane * p.peer_plane[0] + p.peer_off[0];
    if (yl >= p.nyl && p.peer[1]) return p.peer[1] + plane * p.peer_plane[1] + p.peer_off[1];
    return p.A[plane] + rowoff(p, yl);
}

// Source / destination coordinate of a streaming hop with the reference's per-axis periodic wrap
// (src/core/streaming/streaming.cu:13-21).  Returns false when the hop leaves a non-periodic domain.
// yl is a local row; rows -1 and nyl address the ghost rows of a slab interface.
__device__ __forceinline__ bool hop(const Params& p, int& x, int& yl, int dx, int dy) {
    bool ok = true;
    if (dx != 0) {
        x += dx;
        if (x < 0) { if (p.px) x += p.nx; else ok = false; }
        else if (x >= p.nx) { if (p.px) x -= p.nx; else ok = false; }
    }
    if (dy != 0) {
        yl += dy;
        int yg = p.y0 + yl;
        if (yg < 0) { if (p.py) { if (p.wrap_y) yl += p.ny; } else ok = false; }
        else if (yg >= p.ny) { if (p.py) { if (p.wrap_y) yl -= p.ny; } else ok = false; }
    }
    return ok;
}

// index into the edge ring, or -1 for nodes not on a non-periodic domain edge
__device__ __forceinline__ int edge_index(const Params& p, int x, int yg) {
    if (!p.py) { if (yg == 0) return x; if (yg == p.ny - 1) return p.nx + x; }
    if (!p.px) { if (x == 0) return 2 * p.nx + yg; if (x == p.nx - 1) return 2 * p.nx + p.ny + yg; }
    return -1;
}

// Post-stream populations of node (x, yl): step 1 of Appendix D (stream_node, streaming.cu:5-33).
// Slots whose source lies outside a non-periodic domain are "undelivered": the reference leaves the
// node's own post-collision value of two steps ago there (its two-buffer swap), reproduced from the ring.
template <bool ODD>
__device__ __forceinline__ int pull(const Params& p, int x, int yl, float g[Q]) {
    const int gen = p.t & 1;
    const long long row = rowoff(p, yl);
    g[0] = p.A0[gen][row + x];
    if (!ODD) {
#pragma unroll
        for (int q = 1; q < Q; q++) g[q] = p.A[q][row + x];
    } else {
#pragma unroll
        for (int q = 1; q < Q; q++) {
            int xs = x, ys = yl;
            bool ok = hop(p, xs, ys, -cx(q), -cy(q));
            g[q] = ok ? row_ptr(p, opp(q), ys)[xs] : 0.0f;
        }
    }
    const int yg = p.y0 + yl;
    const int e = edge_index(p, x, yg);
    if (e >= 0) {
        const bool ex0 = !p.px && x == 0, ex1 = !p.px && x == p.nx - 1;
        const bool ey0 = !p.py && yg == 0, ey1 = !p.py && yg == p.ny - 1;
        const float* r = p.ring + ((long long)gen * p.perim + e) * Q;
#pragma unroll
        for (int q = 1; q < Q; q++) {
            bool und = (ex0 && cx(q) > 0) || (ex1 && cx(q) < 0) || (ey0 && cy(q) > 0) || (ey1 && cy(q) < 0);
            if (und) g[q] = r[q];
        }
    }
    return e;
}

template <bool ODD>
__device__ __forceinline__ void push(const Params& p, int x, int yl, int e, const float f[Q]) {
    const int gen = p.t & 1;
    const long long row = rowoff(p, yl);
    p.A0[gen][row + x] = f[0];
    if (!ODD) {
#pragma unroll
        for (int q = 1; q < Q; q++) p.A[opp(q)][row + x] = f[q];
    } else {
#pragma unroll
        for (int q = 1; q < Q; q++) {
            int xd = x, yd = yl;
            bool ok = hop(p, xd, yd, cx(q), cy(q));
            if (ok) row_ptr(p, q, yd)[xd] = f[q];
        }
    }
    if (e >= 0) {
        float* r = p.ring + ((long long)gen * p.perim + e) * Q;
#pragma unroll
        for (int q = 0; q < Q; q++) r[q] = f[q];
    }
}

// ------------------------------------------------------------------ equilibrium
// second-order f_eq — reference src/core/equilibrium/equilibrium.cu:5-39 (pure fp32 here)
__device__ __forceinline__ float feq(int q, float rho, float ux, float uy, float usq15) {
    float cu = cx(q) * ux + cy(q) * uy;
    return wq(q) * rho * (1.0f + 3.0f * cu + 4.5f * cu * cu - usq15);
}

struct Moments { float rho, inv_rho, ux, uy, pxx, pxy, pyy; };

// uncorrected_macroscopics_kernel<2> — reference src/core/macroscopics/macroscopics.cu:5-38 (scalar view of moments_v, collide.cuh)
__device__ __forceinline__ Moments moments(const float g[Q]) {
    V1 v[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) v[q].a = g[q];
    const Mom<V1> m = moments_v(v);
    return Moments{m.rho.a, m.inv_rho.a, m.ux.a, m.uy.a, m.pxx.a, m.pxy.a, m.pyy.a};
}

// ------------------------------------------------------------------ boundary functors (Appendix D step 2)
__device__ __forceinline__ int find_sorted(const long long* a, int n, long long key) {
    int lo = 0, hi = n - 1;
    while (lo <= hi) { int mid = (lo + hi) >> 1; long long v = a[mid]; if (v == key) return mid; if (v < key) lo = mid + 1; else hi = mid - 1; }
    return -1;
}

// BounceBack<2>::apply — reference src/functors/boundaryConditions/bbDomainBoundary.cuh:22-49
__device__ __forceinline__ void bc_bounce_back(const Params& p, float g[Q], int x, int yg) {
    const bool raw = (p.quirks & QK_D11) != 0;
#pragma unroll
    for (int i = 1; i < Q; i++) {
        int xn = x + cx(i), yn = yg + cy(i);
        bool xb = (xn < 0 || xn >= p.nx) && (raw || !p.px);
        bool yb = (yn < 0 || yn >= p.ny) && (raw || !p.py);
        if (xb || yb) g[opp(i)] = g[i];
    }
}

// ZouHe::apply_left — reference src/functors/boundaryConditions/zouHeInflow.cuh:9-33
__device__ __forceinline__ void bc_zou_he_left(const Params& p, float g[Q]) {
    const float ux = p.u_max;
    float a = (p.quirks & QK_D3) ? g[2] : g[3];
    float rho = (g[0] + g[2] + g[4] + 2.0f * (a + g[6] + g[7])) / (1.0f - ux);
    g[1] = g[3] + (2.0f / 3.0f) * rho * ux;
    g[5] = g[7] - 0.5f * (g[2] - g[4]) + (1.0f / 6.0f) * rho * ux;
    g[8] = g[6] + 0.5f * (g[2] - g[4]) + (1.0f / 6.0f) * rho * ux;
}

// ZouHe::apply_top — reference zouHeInflow.cuh:36-50 (u = (u_max, 0))
__device__ __forceinline__ void bc_zou_he_top(const Params& p, float g[Q]) {
    const float ux = p.u_max;
    float rho = g[0] + g[1] + g[3] + 2.0f * (g[2] + g[5] + g[6]);
    g[4] = g[2];
    float d13 = g[1] - g[3];
    g[7] = g[5] + 0.5f * d13 - 0.5f * rho * ux;
    g[8] = g[6] - 0.5f * d13 + 0.5f * rho * ux;
}

// CylinderBoundary::apply — reference src/functors/boundaryConditions/cylinderBoundary.cuh:22-34
__device__ __forceinline__ void bc_cylinder(float g[Q]) {
    float t;
    t = g[1]; g[1] = g[3]; g[3] = t;
    t = g[2]; g[2] = g[4]; g[4] = t;
    t = g[5]; g[5] = g[7]; g[7] = t;
    t = g[6]; g[6] = g[8]; g[8] = t;
}

// ZG_OutflowBoundary<2>::apply — reference src/functors/boundaryConditions/zeroGradientOutflow.cuh:9-59
__device__ __forceinline__ void bc_zg_outflow(const Params& p, float g[Q], const float* nb, int x, int yg) {
    int n0 = 0, n1 = 0;
    if (x == 0) n0 = 1; else if (x == p.nx - 1) n0 = -1; else if (yg == 0) n1 = 1; else if (yg == p.ny - 1) n1 = -1; else return;
#pragma unroll
    for (int i = 0; i < Q; i++) if (cx(i) * n0 + cy(i) * n1 > 0) g[i] = nb[i];
}

// PressureOutlet::apply — reference src/functors/boundaryConditions/pressureOutlet.cuh:7-42
__device__ __forceinline__ void bc_pressure_outlet(float g[Q], const float* nb) {
    float h[Q];
#pragma unroll
    for (int i = 0; i < Q; i++) h[i] = nb[i];
    Moments m = moments(h);
    float usq15 = 1.5f * (m.ux * m.ux + m.uy * m.uy);
#pragma unroll
    for (int i = 0; i < Q; i++) g[i] = feq(i, 1.0f, m.ux, m.uy, usq15);
}

// tail shared by RegularizedInlet::apply_top (regularizedInlet.cuh:42-67) and
// RegularizedBounceBack::apply (regularizedBounceBack.cuh:68-96): Pi(1) from all nine f, then f = f_eq + f_neq
__device__ __forceinline__ void regularize(float g[Q], const float fe[Q], float rho, float ux, float uy) {
    const float cs2 = 1.0f / 3.0f;
    float d = g[5] + g[6] + g[7] + g[8];
    float Pxx = g[1] + g[3] + d - (cs2 * rho + rho * ux * ux);
    float Pyy = g[2] + g[4] + d - (cs2 * rho + rho * uy * uy);
    float Pxy = (g[5] - g[6]) + (g[7] - g[8]) - rho * ux * uy;
#pragma unroll
    for (int q = 0; q < Q; q++) {
        float Qxx = cx(q) * cx(q) - cs2, Qyy = cy(q) * cy(q) - cs2, Qxy = (float)(cx(q) * cy(q));
        g[q] = fe[q] + (wq(q) * 4.5f) * (Qxx * Pxx + Qyy * Pyy + 2.0f * Qxy * Pxy);
    }
}

// RegularizedInlet::apply_top — reference src/functors/boundaryConditions/regularizedInlet.cuh:15-69
__device__ __forceinline__ void bc_regularized_inlet_top(const Params& p, float g[Q]) {
    const float ux = p.u_max, uy = 0.0f;
    float rho = g[0] + g[1] + g[3] + 2.0f * (g[2] + g[5] + g[6]);
    float fe[Q];
    float usq15 = 1.5f * ux * ux;
#pragma unroll
    for (int q = 0; q < Q; q++) fe[q] = feq(q, rho, ux, uy, usq15);
#pragma unroll
    for (int q = 0; q < Q; q++) if (cy(q) < 0) g[q] = fe[q] + (g[opp(q)] - fe[opp(q)]);
    regularize(g, fe, rho, ux, uy);
}

// RegularizedBounceBack::apply — reference src/functors/boundaryConditions/regularizedBounceBack.cuh:13-99
__device__ __forceinline__ void bc_regularized_bb(const Params& p, float g[Q], int x, int yg) {
    int sx = 0, sy = 0;       // sign of the unknown directions along the wall normal
    if (x == 0) sx = 1; else if (x == p.nx - 1) sx = -1; else if (yg == 0) sy = 1; else if (yg == p.ny - 1) sy = -1;
    bool unk[Q];
    unk[0] = false;
#pragma unroll
    for (int i = 1; i < Q; i++) unk[i] = (cx(i) * sx + cy(i) * sy) > 0;
    float rho = 0.0f;
#pragma unroll
    for (int i = 0; i < Q; i++) if (!unk[i]) rho += (unk[opp(i)] ? 2.0f : 1.0f) * g[i];
    float fe[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) fe[q] = wq(q) * rho;        // u = 0
#pragma unroll
    for (int q = 0; q < Q; q++) if (unk[q]) g[q] = fe[q] + (g[opp(q)] - fe[opp(q)]);
    regularize(g, fe, rho, 0.0f, 0.0f);
}

// RegularizedCornerBounceBack::apply — reference regularizedBounceBack.cuh:107-221; nb = post-stream
// populations of the diagonal interior node (:136-147)
__device__ __forceinline__ void bc_regularized_corner(const Params& p, float g[Q], const float* nb, int x, int yg) {
    const bool left = x == 0, right = x == p.nx - 1, bottom = yg == 0, top = yg == p.ny - 1;
    if (!((left || right) && (bottom || top))) return;
    float rho = 0.0f;
#pragma unroll
    for (int i = 0; i < Q; i++) rho += nb[i];
#pragma unroll
    for (int i = 1; i < Q; i++) {
        bool unk = (left && cx(i) > 0) || (right && cx(i) < 0) || (bottom && cy(i) > 0) || (top && cy(i) < 0);
        if (unk) g[i] = wq(i) * rho - (g[opp(i)] - wq(opp(i)) * rho);       // minus sign: :164
    }
    // regularize_distributions :192-221 — Pi from f - f_eq (u = 0)
    const float cs2 = 1.0f / 3.0f;
    float n[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) n[q] = g[q] - wq(q) * rho;
    float d = n[5] + n[6] + n[7] + n[8];
    float Pxx = n[1] + n[3] + d, Pyy = n[2] + n[4] + d, Pxy = (n[5] - n[6]) + (n[7] - n[8]);
#pragma unroll
    for (int q = 0; q < Q; q++) {
        float Qxx = cx(q) * cx(q) - cs2, Qyy = cy(q) * cy(q) - cs2, Qxy = (float)(cx(q) * cy(q));
        g[q] = wq(q) * rho + (wq(q) * 4.5f) * (Qxx * Pxx + Qyy * Pyy + 2.0f * Qxy * Pxy);
    }
}

// boundaries_kernel_2D switch — reference src/core/boundaries/boundaries.cuh:29-83
__device__ __forceinline__ void apply_bc(const Params& p, int flag, float g[Q], int x, int yg) {
    switch (flag) {
    case BOUNCE_BACK: bc_bounce_back(p, g, x, yg); break;
    case ZOU_HE_TOP: bc_zou_he_top(p, g); break;
    case ZOU_HE_LEFT: bc_zou_he_left(p, g); break;
    case CYLINDER: bc_cylinder(g); break;
    case REGULARIZED_INLET_TOP: bc_regularized_inlet_top(p, g); break;
    case REGULARIZED_BOUNCE_BACK: bc_regularized_bb(p, g, x, yg); break;
    case ZG_OUTFLOW:
    case PRESSURE_OUTLET:
    case REGULARIZED_BOUNCE_BACK_CORNER: {
        int k = find_sorted(p.nbr_nodes, p.nbr_count, (long long)yg * p.nx + x);
        if (k < 0) break;
        const float* nb = p.nbr_g + (long long)k * Q;
        if (flag == ZG_OUTFLOW) bc_zg_outflow(p, g, nb, x, yg);
        else if (flag == PRESSURE_OUTLET) bc_pressure_outlet(g, nb);
        else bc_regularized_corner(p, g, nb, x, yg);
        break;
    }
    default: break;
    }
}

// collision operators: collide.cuh
__device__ __forceinline__ Relax relax_of(const Params& p) {
    Relax r;
    r.omega = p.omega;
#pragma unroll
    for (int i = 0; i < Q; i++) r.S[i] = p.S[i];
    r.quirks = p.quirks;
    return r;
}
__device__ __forceinline__ AdapterAvg load_adapter_avg(const float* avg) {
    AdapterAvg a; a.inv_rho = fast_rcp(avg[0]); a.inv_j = fast_rcp(avg[1]); a.inv_pi = fast_rcp(avg[2]); return a;
}

}  // namespace lbm
