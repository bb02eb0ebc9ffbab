// kernels.cuh — the __global__ kernels of the engine (see d2q9.cuh for the per-cell functions).
#pragma once
#include "d2q9.cuh"

namespace lbm {

constexpr int BX = 128;     // threads per block along x; one cell per thread, one row per blockIdx.y

// collision ids (include/lbm_b200.h)
constexpr int C_BGK = 0, C_MRT = 1, C_CM = 2, C_CMOPT = 3;

struct NodeState { float g[Q]; Moments m; float Fx, Fy, ux, uy; int e; };

// Steps 1-5 of SURVEY.md Appendix D for one node: pull, boundary functor, moments, force, velocity correction.
template <bool ODD, bool GENERAL>
__device__ __forceinline__ void node_pre_collision(const Params& p, int x, int yl, NodeState& s) {
    s.e = pull<ODD>(p, x, yl, s.g);
    int flag = 0;
    const long long ln = (long long)yl * p.nx + x;
    if (GENERAL) {
        if (p.flags) flag = p.flags[ln];
        const int bc = flag & FLAG_MASK;
        if (bc) apply_bc(p, bc, s.g, x, p.y0 + yl);
    }
    s.m = moments(s.g);
    s.Fx = p.fx; s.Fy = p.fy;                       // reset_forces_kernel -> Init::apply_forces (macroscopics.cuh:13-48)
    if (GENERAL) {
        if (p.force_plane) { float2 F = p.force_plane[ln]; s.Fx = F.x; s.Fy = F.y; }
        if (flag & FLAG_IBM) {                      // IBMManager::multi_direct result (body force already folded in)
            int k = find_sorted(p.ibm_nodes, p.ibm_count, (long long)(p.y0 + yl) * p.nx + x);
            if (k >= 0) { float2 F = p.ibm_force[k]; s.Fx = F.x; s.Fy = F.y; }
        }
    }
    // correct_macroscopics_kernel<2> (macroscopics.cu:99-110): u += F / (2 rho)
    // (0.5f * inv_rho == 0.5f / rho bit for bit: scaling by a power of two; same FMA form as the vector kernel)
    const float h = __fmul_rn(s.m.inv_rho, 0.5f);
    s.ux = s.m.ux; s.uy = s.m.uy;
    if (s.Fx != 0.0f || s.Fy != 0.0f) { s.ux = __fmaf_rn(s.Fx, h, s.m.ux); s.uy = __fmaf_rn(s.Fy, h, s.m.uy); }
}

// Which kernel owns a cell?  The vectorised kernels do stream + collide only: a cell with a boundary flag or under a marker stencil
// (flag byte != 0) or on a non-periodic domain edge (undelivered slots, edge ring) belongs to the general (scalar) kernel.
// Segments (128 cells of a row) are classified once (build_segmask_kernel): SEG_VEC no such cell, SEG_GENERAL all of them (or a
// per-node force plane: everything), SEG_MIXED some — the vectorised kernels then skip exactly those cells, which the general
// kernel takes from a per-cell list.  Under the AA pattern a cell reads exactly the slots it overwrites, so any split is race-free.
constexpr uint8_t SEG_VEC = 0, SEG_GENERAL = 1, SEG_MIXED = 2;
__device__ __forceinline__ bool cell_is_general(const Params& p, int x, int yl) {
    if (p.flags && p.flags[(long long)yl * p.nx + x]) return true;
    if (!p.px && (x == 0 || x == p.nx - 1)) return true;
    const int yg = p.y0 + yl;
    return !p.py && (yg == 0 || yg == p.ny - 1);
}
// the general kernels' three launch shapes: whole slab (grid = segments x rows), listed segments, listed cells
__device__ __forceinline__ bool general_cell_of_thread(const Params& p, int& x, int& yl) {
    if (p.gen_cells) {
        const long long i = (long long)blockIdx.x * BX + threadIdx.x;
        if (i >= p.gen_cell_count) return false;
        const long long ln = p.gen_cells[i];
        yl = (int)(ln / p.nx); x = (int)(ln - (long long)yl * p.nx);
        return true;
    }
    if (p.gen_list) {
        const int seg = p.gen_list[blockIdx.x];
        yl = seg / p.nsx;
        x = (seg - yl * p.nsx) * SEG + threadIdx.x;
    } else {
        x = blockIdx.x * BX + threadIdx.x;
        yl = blockIdx.y;
    }
    return x < p.nx;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block partial sums of (rho, rho|u|, |Pi|): update_avg_mag<2> (macroscopics.cuh:51-120) without atomics
__device__ __forceinline__ void block_partials(float a, float b, float c, float* out) {
    __shared__ float sm[3][BX / 32];
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sm[0][w] = a; sm[1][w] = b; sm[2][w] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < BX / 32; i++)
            if (i < (int)((blockDim.x + 31) >> 5)) { s0 += sm[0][i]; s1 += sm[1][i]; s2 += sm[2][i]; }      // the vector kernel may run narrower blocks
        out[0] = s0; out[1] = s1; out[2] = s2;
    }
}

// the vectorised kernels: one triple per WARP (slot = block * (BX / 32) + warp), no block-wide barrier — a warp retires as soon as its
// own cells are done.  Warps that are switched off (general segment, beyond the row) write zeros.
__device__ __forceinline__ void warp_partials(float a, float b, float c, float* block_out) {
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if ((threadIdx.x & 31) == 0) {
        float* o = block_out + 3 * (threadIdx.x >> 5);
        o[0] = a; o[1] = b; o[2] = c;
    }
}
constexpr int VEC_PARTS = BX / 32;      // partial triples per block of a vectorised kernel

// The fused step: one launch = one reference time step (src/main.cu:96-114) for every node it covers.
// Scalar form, one cell per thread.  Two launch shapes: the whole slab (grid = segments x rows; used when nx is not a
// multiple of 4) or only the "general" segments listed in p.gen_list (grid.x = number of listed segments), the
// vectorised kernel below covering everything else.
template <int COLL, bool ODD, bool GENERAL>
__global__ void __launch_bounds__(BX) step_kernel(const Params p) {
    int x = 0, yl = 0;
    const bool on = general_cell_of_thread(p, x, yl);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (on) {
        NodeState s;
        node_pre_collision<ODD, GENERAL>(p, x, yl, s);
        if (p.rho_out) {
            const long long ln = (long long)yl * p.nx + x;
            p.rho_out[ln] = s.m.rho;
            p.u_out[ln] = make_float2(s.ux, s.uy);
        }
        V1 gv[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) gv[q].a = s.g[q];
        const V1 rho{s.m.rho}, ux{s.ux}, uy{s.uy}, Fx{s.Fx}, Fy{s.Fy};
        const bool forced = s.Fx != 0.0f || s.Fy != 0.0f;
        const Relax rx = relax_of(p);
        if (COLL == C_BGK) collide_bgk_v(rx, gv, rho, ux, uy, forced, Fx, Fy);
        else if (COLL == C_MRT) collide_mrt_v(rx, gv, rho, ux, uy, forced, Fx, Fy);
        else if (COLL == C_CM) collide_cm_v<false>(rx, gv, ux, uy, forced, Fx, Fy, V1{1.0f});
        else {
            const V1 jm = jmag_v(ux, uy, rho), pm = pi_norm_v(Mom<V1>{rho, V1{s.m.inv_rho}, V1{s.m.ux}, V1{s.m.uy}, V1{s.m.pxx}, V1{s.m.pxy}, V1{s.m.pyy}});
            s0 = s.m.rho; s1 = jm.a; s2 = pm.a;
            const AdapterAvg av = load_adapter_avg(p.avg);
            collide_cm_v<true>(rx, gv, ux, uy, forced, Fx, Fy, optimal_rate_v(rho, jm, pm, av));
        }
#pragma unroll
        for (int q = 0; q < Q; q++) s.g[q] = gv[q].a;
        push<ODD>(p, x, yl, s.e, s.g);
    }
    if (COLL == C_CMOPT && p.partials) block_partials(s0, s1, s2, p.partials + 3 * ((long long)blockIdx.y * gridDim.x + blockIdx.x));
}

// ------------------------------------------------------------------ vectorised fast path
// Four consecutive cells of one row per thread, 128-bit loads and stores on every slot plane.  Covers the cells that
// need nothing but stream + collide (FLUID, uniform body force, away from non-periodic domain edges): everything of a
// periodic Taylor-Green box, all but O(perimeter + bodies) segments elsewhere.
//   even step: every slot is read and written at the cell's own address -> aligned float4.
//   odd step : slot planes with c_x != 0 are shifted by one cell.  Each thread still issues one ALIGNED float4 access per
//              plane and the one element that crosses the 16-byte boundary moves between neighbouring lanes with a
//              warp shuffle; only the first / last lane of a warp (or of a row) touch the odd element with a scalar access.
constexpr unsigned FULL = 0xffffffffu;
// 128-thread blocks per SM the register allocator must allow (even / odd AA phase)
#ifndef LBM_VEC_MIN_BLOCKS_EVEN
#define LBM_VEC_MIN_BLOCKS_EVEN 6
#endif
#ifndef LBM_VEC_MIN_BLOCKS_ODD
#define LBM_VEC_MIN_BLOCKS_ODD 4     // 128 registers: the odd phase keeps nine float4 loads in flight next to the 36 population registers
#endif
// CM<2,OptimalAdapter> carries the adapter quantities and partial sums on top
#ifndef LBM_VEC_MIN_BLOCKS_EVEN_OPT
#define LBM_VEC_MIN_BLOCKS_EVEN_OPT 5
#endif
#ifndef LBM_VEC_MIN_BLOCKS_ODD_OPT
#define LBM_VEC_MIN_BLOCKS_ODD_OPT 4
#endif
constexpr int vec_min_blocks(int coll, bool odd) {
    return coll == 3 ? (odd ? LBM_VEC_MIN_BLOCKS_ODD_OPT : LBM_VEC_MIN_BLOCKS_EVEN_OPT) : (odd ? LBM_VEC_MIN_BLOCKS_ODD : LBM_VEC_MIN_BLOCKS_EVEN);
}

// cache hints of the population accesses (every population is read once and written once per step):
// LBM_LD_HINT 0 plain, 1 ld.global.cs (streaming), 2 ld.global.lu (last use); LBM_ST_HINT 0 plain, 1 st.global.cs, 2 st.global.wt
#ifndef LBM_LD_HINT
#define LBM_LD_HINT 0
#endif
#ifndef LBM_AVG_EARLY
#define LBM_AVG_EARLY 1
#endif
#ifndef LBM_ST_HINT
#define LBM_ST_HINT 0
#endif
__device__ __forceinline__ float4 ld4(const float* p) {
#if LBM_LD_HINT == 1
    return __ldcs(reinterpret_cast<const float4*>(p));
#elif LBM_LD_HINT == 2
    return __ldlu(reinterpret_cast<const float4*>(p));
#else
    return *reinterpret_cast<const float4*>(p);
#endif
}
__device__ __forceinline__ void st4(float* p, float a, float b, float c, float d) {
#if LBM_ST_HINT == 1
    __stcs(reinterpret_cast<float4*>(p), make_float4(a, b, c, d));
#elif LBM_ST_HINT == 2
    __stwt(reinterpret_cast<float4*>(p), make_float4(a, b, c, d));
#else
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
#endif
}

// class of segment sx of row yl; no memory access inside the all-vector rectangle
__device__ __forceinline__ uint8_t segment_class(const Params& p, int yl, int sx) {
    if (!p.segmask) return SEG_VEC;
    if (yl >= p.pure_y0 && yl < p.pure_y1) {
        if (sx >= p.pure_s0 && sx < p.pure_s1) return SEG_VEC;
        if (sx < COLCLASS_MAX) { const uint8_t c = p.colclass[sx]; if (c != SEG_LOOKUP) return c; }     // kernel parameter space: no memory round trip
    }
    return p.segmask[(long long)yl * p.nsx + sx];
}

// SEG_MIXED segments: bit k set = cell x0 + k of this thread belongs to the general kernel
__device__ __forceinline__ unsigned skip_mask4(const Params& p, int x0, int yl) {
    unsigned m = 0;
    if (p.flags) {
        const uchar4 f = *reinterpret_cast<const uchar4*>(p.flags + (long long)yl * p.nx + x0);
        m = (f.x ? 1u : 0u) | (f.y ? 2u : 0u) | (f.z ? 4u : 0u) | (f.w ? 8u : 0u);
    }
    if (!p.px) { if (x0 == 0) m |= 1u; if (x0 + 4 == p.nx) m |= 8u; }
    return m;
}
__device__ __forceinline__ void st4_masked(float* ptr, float a, float b, float c, float d, unsigned skip) {
    if (skip == 0) { st4(ptr, a, b, c, d); return; }
    if (!(skip & 1u)) ptr[0] = a;
    if (!(skip & 2u)) ptr[1] = b;
    if (!(skip & 4u)) ptr[2] = c;
    if (!(skip & 8u)) ptr[3] = d;
}
__device__ __forceinline__ V2 zero_lanes(V2 v, bool z0, bool z1) { V2 r; r.a = make_float2(z0 ? 0.0f : v.a.x, z1 ? 0.0f : v.a.y); return r; }

// f[0..3] of cells x0..x0+3 go to x0+1 .. x0+4
__device__ __forceinline__ void store_to_right(float* r, int x0, int xr, bool hasL, bool hasR, bool act, const float f[4]) {
    const float l = __shfl_up_sync(FULL, f[3], 1);
    if (!act) return;
    if (hasL) st4(r + x0, l, f[0], f[1], f[2]);
    else { r[x0 + 1] = f[0]; r[x0 + 2] = f[1]; r[x0 + 3] = f[2]; }
    if (!hasR) r[xr] = f[3];
}
// f[0..3] go to x0-1 .. x0+2
__device__ __forceinline__ void store_to_left(float* r, int x0, int xl, bool hasL, bool hasR, bool act, const float f[4]) {
    const float rr = __shfl_down_sync(FULL, f[0], 1);
    if (!act) return;
    if (hasR) st4(r + x0, f[1], f[2], f[3], rr);
    else { r[x0] = f[1]; r[x0 + 1] = f[2]; r[x0 + 2] = f[3]; }
    if (!hasL) r[xl] = f[0];
}

// Where a thread of the vectorised kernels works: cells x0 .. x0+3 of row yl, and how it reaches rows y-1 / y+1.
struct VecCtx {
    int yl, x0, xl, xr, gen;
    unsigned skip;              // SEG_MIXED: cells of this thread the general kernel owns
    bool act, hasL, hasR;
    long long r0;
    float *bm, *bp;             // rows y-1 / y+1 of plane k start at bm + k*sm / bp + k*sp: this slab, its ghost rows, or (peer-mapped) the neighbour slab
    long long sm, sp;
};

// returns false for warps that have nothing to do (beyond the row end, or a segment the general kernel owns)
template <bool ODD>
__device__ __forceinline__ bool vec_setup(const Params& p, VecCtx& c) {
    const int nv = p.nx >> 2;
    const int xv_raw = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    c.yl = blockIdx.y;
    c.act = xv_raw < nv;
    const int xv = c.act ? xv_raw : nv - 1;           // idle lanes of the last warp shadow a valid cell (loads only)
    bool warp_on = (xv_raw & ~31) < nv;               // warp-uniform: one warp = one 128-cell segment
    uint8_t segtype = SEG_VEC;
    if (warp_on) { segtype = segment_class(p, c.yl, xv_raw >> 5); warp_on = segtype != SEG_GENERAL; }
    if (!warp_on) return false;
    c.x0 = xv << 2;
    c.skip = segtype == SEG_MIXED ? skip_mask4(p, c.x0, c.yl) : 0u;
    c.r0 = rowoff(p, c.yl);
    c.gen = p.t & 1;
    // neighbours inside the warp exchange the boundary element; a row starts at lane 0 (blockDim.x % 32 == 0)
    c.hasL = lane > 0; c.hasR = lane < 31 && xv_raw < nv - 1;
    c.xl = c.x0 - 1; c.xr = c.x0 + 4;
    if (c.xl < 0) c.xl += p.nx;
    if (c.xr >= p.nx) c.xr -= p.nx;
    int ym = c.yl - 1, yp = c.yl + 1;
    if (p.wrap_y) { if (ym < 0) ym += p.nyl; if (yp >= p.nyl) yp -= p.nyl; }
    c.bm = p.A[0] + rowoff(p, ym); c.sm = p.plane;
    c.bp = p.A[0] + rowoff(p, yp); c.sp = p.plane;
    if (ODD) {
        if (ym < 0 && p.peer[0]) { c.bm = p.peer[0] + p.peer_off[0]; c.sm = p.peer_plane[0]; }
        if (yp >= p.nyl && p.peer[1]) { c.bp = p.peer[1] + p.peer_off[1]; c.sp = p.peer_plane[1]; }
    }
    return true;
}

// post-stream populations of the thread's four cells: g[h][q] holds cells x0+2h, x0+2h+1 side by side (packed fp32 lanes)
template <bool ODD>
__device__ __forceinline__ void vec_load(const Params& p, const VecCtx& c, V2 g[2][Q]) {
    const int x0 = c.x0;
    {
        float4 v = ld4(p.A0[c.gen] + c.r0 + x0);
        g[0][0].a = make_float2(v.x, v.y); g[1][0].a = make_float2(v.z, v.w);
    }
    if (!ODD) {
#pragma unroll
        for (int q = 1; q < Q; q++) {
            float4 v = ld4(p.A[q] + c.r0 + x0);
            g[0][q].a = make_float2(v.x, v.y); g[1][q].a = make_float2(v.z, v.w);
        }
    } else {
        // g_q(x) = A[opp q][x - c_q]: source row y - c_y, source column x - c_x.  All vector loads first, then the
        // (predicated) scalar loads of the first / last lane, then the shuffles -- nothing between the loads that
        // would make them wait for one another.
        float4 v[Q]; float e[Q];
#pragma unroll
        for (int q = 1; q < Q; q++) {
            const float* r = cy(q) > 0 ? c.bm + opp(q) * c.sm : (cy(q) < 0 ? c.bp + opp(q) * c.sp : p.A[opp(q)] + c.r0);
            v[q] = ld4(r + x0);
            e[q] = 0.f;
            if (cx(q) > 0) { if (!c.hasL) e[q] = r[c.xl]; }
            else if (cx(q) < 0) { if (!c.hasR) e[q] = r[c.xr]; }
        }
#pragma unroll
        for (int q = 1; q < Q; q++) {
            if (cx(q) > 0) {
                float l = __shfl_up_sync(FULL, v[q].w, 1);
                if (!c.hasL) l = e[q];
                g[0][q].a = make_float2(l, v[q].x); g[1][q].a = make_float2(v[q].y, v[q].z);
            } else if (cx(q) < 0) {
                float rr = __shfl_down_sync(FULL, v[q].x, 1);
                if (!c.hasR) rr = e[q];
                g[0][q].a = make_float2(v[q].y, v[q].z); g[1][q].a = make_float2(v[q].w, rr);
            } else { g[0][q].a = make_float2(v[q].x, v[q].y); g[1][q].a = make_float2(v[q].z, v[q].w); }
        }
    }
}

// NOTE: setup and loads are written out inside this kernel on purpose.  Routing them through vec_setup / vec_load (as the
// moments pre-pass does) is semantically identical but changed nvcc's schedule enough to cost 12 % on the odd phase
// (5.80 vs 6.58 TB/s at 16384^2, profiles/r01_kbench_packed.txt), with the same register count.
template <int COLL, bool ODD>
__global__ void __launch_bounds__(BX, vec_min_blocks(COLL, ODD)) step_vec_kernel(const Params p) {
    const int nv = p.nx >> 2;
    const int xv_raw = blockIdx.x * blockDim.x + threadIdx.x;
    const int yl = blockIdx.y;
    const int lane = threadIdx.x & 31;
    bool act = xv_raw < nv;
    const int xv = act ? xv_raw : nv - 1;             // idle lanes of the last warp shadow a valid cell (loads only)
    bool warp_on = (xv_raw & ~31) < nv;               // warp-uniform: one warp = one 128-cell segment
    uint8_t segtype = SEG_VEC;
    if (warp_on) { segtype = segment_class(p, yl, xv_raw >> 5); warp_on = segtype != SEG_GENERAL; }
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (warp_on) {
        const int x0 = xv << 2;
        // cells of this thread the general kernel owns (even phase only: the host never hands SEG_MIXED segments to the odd branch of
        // this kernel — step_odd_kernel below is the odd phase wherever segments can be mixed)
        const unsigned skip = (!ODD && segtype == SEG_MIXED) ? skip_mask4(p, x0, yl) : 0u;
        const long long r0 = rowoff(p, yl);
        const int gen = p.t & 1;
        // CM<2,OptimalAdapter>: the three grid means come from memory (the previous launch wrote them).  Their loads are issued
        // here, in front of the population loads, so that all of them are in flight together — behind the populations they
        // would cost a second full memory latency per thread.
        float avg_raw[3] = {1.f, 1.f, 1.f};
#if LBM_AVG_EARLY
        if (COLL == C_CMOPT) { avg_raw[0] = __ldg(p.avg); avg_raw[1] = __ldg(p.avg + 1); avg_raw[2] = __ldg(p.avg + 2); }
#endif
        V2 g[2][Q];                                   // g[h][q]: cells x0+2h, x0+2h+1 side by side (packed fp32 lanes)
        // neighbours inside the warp exchange the boundary element; a row starts at lane 0 (blockDim.x % 32 == 0)
        const bool hasL = lane > 0, hasR = lane < 31 && xv_raw < nv - 1;
        int xl = x0 - 1, xr = x0 + 4;
        if (xl < 0) xl += p.nx;
        if (xr >= p.nx) xr -= p.nx;
        int ym = yl - 1, yp = yl + 1;
        if (p.wrap_y) { if (ym < 0) ym += p.nyl; if (yp >= p.nyl) yp -= p.nyl; }
        // rows y-1 / y+1 of plane k start at bm + k*sm / bp + k*sp: this slab, its ghost rows, or (peer-mapped) the neighbour slab
        float* bm = p.A[0] + rowoff(p, ym); long long sm = p.plane;
        float* bp = p.A[0] + rowoff(p, yp); long long sp = p.plane;
        if (ODD) {
            if (ym < 0 && p.peer[0]) { bm = p.peer[0] + p.peer_off[0]; sm = p.peer_plane[0]; }
            if (yp >= p.nyl && p.peer[1]) { bp = p.peer[1] + p.peer_off[1]; sp = p.peer_plane[1]; }
        }
        {
            float4 v = ld4(p.A0[gen] + r0 + x0);
            g[0][0].a = make_float2(v.x, v.y); g[1][0].a = make_float2(v.z, v.w);
        }
        if (!ODD) {
#pragma unroll
            for (int q = 1; q < Q; q++) {
                float4 v = ld4(p.A[q] + r0 + x0);
                g[0][q].a = make_float2(v.x, v.y); g[1][q].a = make_float2(v.z, v.w);
            }
        } else {
            // g_q(x) = A[opp q][x - c_q]: source row y - c_y, source column x - c_x.  All vector loads first, then the
            // (predicated) scalar loads of the first / last lane, then the shuffles -- nothing between the loads that
            // would make them wait for one another.
            float4 v[Q]; float e[Q];
#pragma unroll
            for (int q = 1; q < Q; q++) {
                const float* r = cy(q) > 0 ? bm + opp(q) * sm : (cy(q) < 0 ? bp + opp(q) * sp : p.A[opp(q)] + r0);
                v[q] = ld4(r + x0);
                e[q] = 0.f;
                if (cx(q) > 0) { if (!hasL) e[q] = r[xl]; }
                else if (cx(q) < 0) { if (!hasR) e[q] = r[xr]; }
            }
#pragma unroll
            for (int q = 1; q < Q; q++) {
                if (cx(q) > 0) {
                    float l = __shfl_up_sync(FULL, v[q].w, 1);
                    if (!hasL) l = e[q];
                    g[0][q].a = make_float2(l, v[q].x); g[1][q].a = make_float2(v[q].y, v[q].z);
                } else if (cx(q) < 0) {
                    float rr = __shfl_down_sync(FULL, v[q].x, 1);
                    if (!hasR) rr = e[q];
                    g[0][q].a = make_float2(v[q].y, v[q].z); g[1][q].a = make_float2(v[q].w, rr);
                } else { g[0][q].a = make_float2(v[q].x, v[q].y); g[1][q].a = make_float2(v[q].z, v[q].w); }
            }
        }
        float2 rho4[2], ux4[2], uy4[2];
        AdapterAvg av{};
#if LBM_AVG_EARLY
        if (COLL == C_CMOPT) { av.inv_rho = fast_rcp(avg_raw[0]); av.inv_j = fast_rcp(avg_raw[1]); av.inv_pi = fast_rcp(avg_raw[2]); }
#else
        if (COLL == C_CMOPT) av = load_adapter_avg(p.avg);
#endif
        const Relax rx = relax_of(p);
        const bool forced = p.fx != 0.0f || p.fy != 0.0f;      // uniform body force only: everything else takes the general path
        const V2 Fx = splat<V2>(p.fx), Fy = splat<V2>(p.fy);
        V2 acc0 = splat<V2>(0.f), acc1 = acc0, acc2 = acc0;
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
            const Mom<V2> m = moments_v(g[hf]);
            // correct_macroscopics_kernel<2> (macroscopics.cu:99-110): u += F / (2 rho)
            V2 ux = m.ux, uy = m.uy;
            if (forced) { const V2 hr = m.inv_rho * 0.5f; ux = fma(Fx, hr, ux); uy = fma(Fy, hr, uy); }
            rho4[hf] = m.rho.a; ux4[hf] = ux.a; uy4[hf] = uy.a;
            if (COLL == C_BGK) collide_bgk_v(rx, g[hf], m.rho, ux, uy, forced, Fx, Fy);
            else if (COLL == C_MRT) collide_mrt_v(rx, g[hf], m.rho, ux, uy, forced, Fx, Fy);
            else if (COLL == C_CM) collide_cm_v<false>(rx, g[hf], ux, uy, forced, Fx, Fy, splat<V2>(1.0f));
            else {
                const V2 jm = jmag_v(ux, uy, m.rho), pm = pi_norm_v(m);
                const bool z0 = (skip >> (2 * hf)) & 1u, z1 = (skip >> (2 * hf + 1)) & 1u;        // their sums come from the general kernel
                acc0 = acc0 + zero_lanes(m.rho, z0, z1); acc1 = acc1 + zero_lanes(jm, z0, z1); acc2 = acc2 + zero_lanes(pm, z0, z1);
                collide_cm_v<true>(rx, g[hf], ux, uy, forced, Fx, Fy, optimal_rate_v(m.rho, jm, pm, av));
            }
        }
        if (COLL == C_CMOPT && act) { s0 = hsum(acc0); s1 = hsum(acc1); s2 = hsum(acc2); }
        if (p.rho_out && act) {
            const long long ln = (long long)yl * p.nx + x0;
            st4_masked(p.rho_out + ln, rho4[0].x, rho4[0].y, rho4[1].x, rho4[1].y, skip);
            float* uo = reinterpret_cast<float*>(p.u_out + ln);
            st4_masked(uo, ux4[0].x, uy4[0].x, ux4[0].y, uy4[0].y, (skip & 1u ? 3u : 0u) | (skip & 2u ? 12u : 0u));
            st4_masked(uo + 4, ux4[1].x, uy4[1].x, ux4[1].y, uy4[1].y, (skip & 4u ? 3u : 0u) | (skip & 8u ? 12u : 0u));
        }
        if (act) st4_masked(p.A0[gen] + r0 + x0, g[0][0].a.x, g[0][0].a.y, g[1][0].a.x, g[1][0].a.y, skip);
        if (!ODD) {
            if (act) {
#pragma unroll
                for (int q = 1; q < Q; q++) st4_masked(p.A[opp(q)] + r0 + x0, g[0][q].a.x, g[0][q].a.y, g[1][q].a.x, g[1][q].a.y, skip);
            }
        } else {
#pragma unroll
            for (int q = 1; q < Q; q++) {
                // f*_q(x) -> A[q][x + c_q]: destination row y + c_y, destination column x + c_x
                float* r = cy(q) > 0 ? bp + q * sp : (cy(q) < 0 ? bm + q * sm : p.A[q] + r0);
                const float f[4] = {g[0][q].a.x, g[0][q].a.y, g[1][q].a.x, g[1][q].a.y};
                if (cx(q) > 0) store_to_right(r, x0, xr, hasL, hasR, act, f);
                else if (cx(q) < 0) store_to_left(r, x0, xl, hasL, hasR, act, f);
                else if (act) st4(r + x0, f[0], f[1], f[2], f[3]);
            }
        }
    }
    if (COLL == C_CMOPT && p.partials) warp_partials(s0, s1, s2, p.partials + 3 * VEC_PARTS * ((long long)blockIdx.y * gridDim.x + blockIdx.x));
}

// ------------------------------------------------------------------ odd (neighbour) phase, lane-interleaved cells
// The odd AA phase reads A[opp q][x - c_q] and writes A[q][x + c_q]: six of the nine slot planes are shifted by one cell in x.  The
// kernel above keeps 128-bit accesses for them (aligned float4 + lane shuffles + predicated scalar accesses for the element that
// crosses a 16-byte boundary), which ncu shows to cost ~280 of its ~1000 warp instructions per 128 cells, 76 of them register moves
// that re-pair the shifted elements for the packed fp32 instructions (profiles/r02_ncu_cm_opt_odd_before.md), and 36 staging registers.
// (TMA cannot take the shift either: a tensor-map box must start on a 16-byte boundary in global memory — tools/tma_probe.cu,
// profiles/r02_tma_probe.txt.)  Here a warp still owns one 128-cell segment, but lane l works on cells x0 + l, l + 32, l + 64, l + 96:
// every access is a 32-bit load / store that the warp coalesces into one 128-byte request whatever the shift, the two cells of a
// packed pair are loaded straight into a register pair, and there is nothing to shuffle, predicate or re-pair.  Same per-cell
// arithmetic as every other kernel (bit-identical results).  Needs rows made of whole segments (nx % 128 == 0).
#ifndef LBM_ODD_MIN_BLOCKS
#define LBM_ODD_MIN_BLOCKS 4      // 5 (<= 102 registers) spills 16 bytes and loses 1.3 % (profiles/r02_kbench_odd_kernel.txt)
#endif
#ifndef LBM_ODD_MIN_BLOCKS_OPT
#define LBM_ODD_MIN_BLOCKS_OPT 4
#endif
template <int COLL>
__global__ void __launch_bounds__(BX, COLL == 3 ? LBM_ODD_MIN_BLOCKS_OPT : LBM_ODD_MIN_BLOCKS) step_odd_kernel(const Params p) {
    const int lane = threadIdx.x & 31;
    const int sx = blockIdx.x * (BX / 32) + (threadIdx.x >> 5);       // segment of this warp within the row
    const int yl = blockIdx.y;
    bool warp_on = sx < p.nsx;
    uint8_t segtype = SEG_VEC;
    if (warp_on) { segtype = segment_class(p, yl, sx); warp_on = segtype != SEG_GENERAL; }
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (warp_on) {
        float avg_raw[3] = {1.f, 1.f, 1.f};
        if (COLL == C_CMOPT) { avg_raw[0] = __ldg(p.avg); avg_raw[1] = __ldg(p.avg + 1); avg_raw[2] = __ldg(p.avg + 2); }
        const int x = sx * SEG + lane;                                // cells x, x + 32, x + 64, x + 96
        const long long r0 = rowoff(p, yl);
        const int gen = p.t & 1;
        int ym = yl - 1, yp = yl + 1;
        if (p.wrap_y) { if (ym < 0) ym += p.nyl; if (yp >= p.nyl) yp -= p.nyl; }
        // rows y-1 / y+1 of plane k start at bm + k*sm / bp + k*sp: this slab, its ghost rows, or (peer-mapped) the neighbour slab
        float* bm = p.A[0] + rowoff(p, ym); long long sm = p.plane;
        float* bp = p.A[0] + rowoff(p, yp); long long sp = p.plane;
        if (ym < 0 && p.peer[0]) { bm = p.peer[0] + p.peer_off[0]; sm = p.peer_plane[0]; }
        if (yp >= p.nyl && p.peer[1]) { bp = p.peer[1] + p.peer_off[1]; sp = p.peer_plane[1]; }
        // the one element per row end that wraps around periodically (a non-periodic row end is a general cell, never computed here)
        const int xw_lo = x == 0 ? p.nx - 1 : x - 1;                  // source column of cell x for c_x = +1, destination for c_x = -1
        const int xw_hi = x + 97 == p.nx ? 0 : x + 97;                // source column of cell x + 96 for c_x = -1, destination for c_x = +1
        unsigned skip = 0;                                            // SEG_MIXED: bit j = cell x + 32 j belongs to the general kernel
        if (segtype == SEG_MIXED) {
            if (p.flags) {
                const uint8_t* f = p.flags + (long long)yl * p.nx + x;
                skip = (f[0] ? 1u : 0u) | (f[32] ? 2u : 0u) | (f[64] ? 4u : 0u) | (f[96] ? 8u : 0u);
            }
            if (!p.px) { if (x == 0) skip |= 1u; if (x + 97 == p.nx) skip |= 8u; }
        }
        V2 g[2][Q];                                                   // g[h][q]: cells x + 64 h and x + 64 h + 32 in the two packed fp32 lanes
        {
            const float* r = p.A0[gen] + r0 + x;
            g[0][0].a.x = r[0]; g[0][0].a.y = r[32]; g[1][0].a.x = r[64]; g[1][0].a.y = r[96];
        }
#pragma unroll
        for (int q = 1; q < Q; q++) {
            // g_q(x) = A[opp q][x - c_q]: source row y - c_y, source column x - c_x
            const float* r = cy(q) > 0 ? bm + opp(q) * sm : (cy(q) < 0 ? bp + opp(q) * sp : p.A[opp(q)] + r0);
            if (cx(q) > 0) { g[0][q].a.x = r[xw_lo]; g[0][q].a.y = r[x + 31]; g[1][q].a.x = r[x + 63]; g[1][q].a.y = r[x + 95]; }
            else if (cx(q) < 0) { g[0][q].a.x = r[x + 1]; g[0][q].a.y = r[x + 33]; g[1][q].a.x = r[x + 65]; g[1][q].a.y = r[xw_hi]; }
            else { g[0][q].a.x = r[x]; g[0][q].a.y = r[x + 32]; g[1][q].a.x = r[x + 64]; g[1][q].a.y = r[x + 96]; }
        }
        float2 rho4[2], ux4[2], uy4[2];
        AdapterAvg av{};
        if (COLL == C_CMOPT) { av.inv_rho = fast_rcp(avg_raw[0]); av.inv_j = fast_rcp(avg_raw[1]); av.inv_pi = fast_rcp(avg_raw[2]); }
        const Relax rx = relax_of(p);
        const bool forced = p.fx != 0.0f || p.fy != 0.0f;
        const V2 Fx = splat<V2>(p.fx), Fy = splat<V2>(p.fy);
        V2 acc0 = splat<V2>(0.f), acc1 = acc0, acc2 = acc0;
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
            const Mom<V2> m = moments_v(g[hf]);
            V2 ux = m.ux, uy = m.uy;
            if (forced) { const V2 hr = m.inv_rho * 0.5f; ux = fma(Fx, hr, ux); uy = fma(Fy, hr, uy); }
            rho4[hf] = m.rho.a; ux4[hf] = ux.a; uy4[hf] = uy.a;
            if (COLL == C_BGK) collide_bgk_v(rx, g[hf], m.rho, ux, uy, forced, Fx, Fy);
            else if (COLL == C_MRT) collide_mrt_v(rx, g[hf], m.rho, ux, uy, forced, Fx, Fy);
            else if (COLL == C_CM) collide_cm_v<false>(rx, g[hf], ux, uy, forced, Fx, Fy, splat<V2>(1.0f));
            else {
                const V2 jm = jmag_v(ux, uy, m.rho), pm = pi_norm_v(m);
                const bool z0 = (skip >> (2 * hf)) & 1u, z1 = (skip >> (2 * hf + 1)) & 1u;
                acc0 = acc0 + zero_lanes(m.rho, z0, z1); acc1 = acc1 + zero_lanes(jm, z0, z1); acc2 = acc2 + zero_lanes(pm, z0, z1);
                collide_cm_v<true>(rx, g[hf], ux, uy, forced, Fx, Fy, optimal_rate_v(m.rho, jm, pm, av));
            }
        }
        if (COLL == C_CMOPT) { s0 = hsum(acc0); s1 = hsum(acc1); s2 = hsum(acc2); }
        const bool k0 = !(skip & 1u), k1 = !(skip & 2u), k2 = !(skip & 4u), k3 = !(skip & 8u);
        if (p.rho_out) {
            const long long ln = (long long)yl * p.nx + x;
            if (k0) { p.rho_out[ln] = rho4[0].x; p.u_out[ln] = make_float2(ux4[0].x, uy4[0].x); }
            if (k1) { p.rho_out[ln + 32] = rho4[0].y; p.u_out[ln + 32] = make_float2(ux4[0].y, uy4[0].y); }
            if (k2) { p.rho_out[ln + 64] = rho4[1].x; p.u_out[ln + 64] = make_float2(ux4[1].x, uy4[1].x); }
            if (k3) { p.rho_out[ln + 96] = rho4[1].y; p.u_out[ln + 96] = make_float2(ux4[1].y, uy4[1].y); }
        }
        {
            float* r = p.A0[gen] + r0 + x;
            if (k0) r[0] = g[0][0].a.x;
            if (k1) r[32] = g[0][0].a.y;
            if (k2) r[64] = g[1][0].a.x;
            if (k3) r[96] = g[1][0].a.y;
        }
#pragma unroll
        for (int q = 1; q < Q; q++) {
            // f*_q(x) -> A[q][x + c_q]: destination row y + c_y, destination column x + c_x
            float* r = cy(q) > 0 ? bp + q * sp : (cy(q) < 0 ? bm + q * sm : p.A[q] + r0);
            const int d0 = cx(q) > 0 ? x + 1 : (cx(q) < 0 ? xw_lo : x);
            const int d3 = cx(q) > 0 ? xw_hi : (cx(q) < 0 ? x + 95 : x + 96);
            if (k0) r[d0] = g[0][q].a.x;
            if (k1) r[x + 32 + cx(q)] = g[0][q].a.y;
            if (k2) r[x + 64 + cx(q)] = g[1][q].a.x;
            if (k3) r[d3] = g[1][q].a.y;
        }
    }
    if (COLL == C_CMOPT && p.partials) warp_partials(s0, s1, s2, p.partials + 3 * VEC_PARTS * ((long long)blockIdx.y * gridDim.x + blockIdx.x));
}

// moments pre-pass for LBM_ADAPTER_EXACT: the grid sums of the CURRENT post-stream state, before any cell collides.
// Same two launch shapes as the step: this scalar kernel over the whole slab or over the listed general segments ...
template <bool ODD>
__global__ void __launch_bounds__(BX) moments_kernel(const Params p) {
    int x = 0, yl = 0;
    const bool on = general_cell_of_thread(p, x, yl);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (on) {
        NodeState s;
        node_pre_collision<ODD, true>(p, x, yl, s);
        const V1 rho{s.m.rho};
        s0 = s.m.rho; s1 = jmag_v(V1{s.ux}, V1{s.uy}, rho).a;
        s2 = pi_norm_v(Mom<V1>{rho, V1{s.m.inv_rho}, V1{s.m.ux}, V1{s.m.uy}, V1{s.m.pxx}, V1{s.m.pxy}, V1{s.m.pyy}}).a;
    }
    block_partials(s0, s1, s2, p.partials + 3 * ((long long)blockIdx.y * gridDim.x + blockIdx.x));
}
// ... and the vectorised one over everything else (36 B/cell read, nothing written)
template <bool ODD>
__global__ void __launch_bounds__(BX, 6) moments_vec_kernel(const Params p) {
    VecCtx c;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (vec_setup<ODD>(p, c)) {
        V2 g[2][Q];
        vec_load<ODD>(p, c, g);
        const bool forced = p.fx != 0.0f || p.fy != 0.0f;
        V2 acc0 = splat<V2>(0.f), acc1 = acc0, acc2 = acc0;
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
            const Mom<V2> m = moments_v(g[hf]);
            V2 ux = m.ux, uy = m.uy;
            if (forced) { const V2 hr = m.inv_rho * 0.5f; ux = fma(splat<V2>(p.fx), hr, ux); uy = fma(splat<V2>(p.fy), hr, uy); }
            const bool z0 = (c.skip >> (2 * hf)) & 1u, z1 = (c.skip >> (2 * hf + 1)) & 1u;
            acc0 = acc0 + zero_lanes(m.rho, z0, z1); acc1 = acc1 + zero_lanes(jmag_v(ux, uy, m.rho), z0, z1); acc2 = acc2 + zero_lanes(pi_norm_v(m), z0, z1);
        }
        if (c.act) { s0 = hsum(acc0); s1 = hsum(acc1); s2 = hsum(acc2); }
    }
    warp_partials(s0, s1, s2, p.partials + 3 * VEC_PARTS * ((long long)blockIdx.y * gridDim.x + blockIdx.x));
}

// deterministic two-level reduction of the block partials: stage 1 (many blocks) folds the fp32 partials into
// at most RED_BLOCKS fp64 triples, stage 2 (one block) produces sums[3] and avg[3] = sums / (NX*NY)
constexpr int RED_BLOCKS = 592;     // 4 per SM
template <typename T>
__device__ __forceinline__ void block_sum3(const T* in, long long n, long long start, long long stride, double out[3]) {
    __shared__ double sm[3][256];
    double a = 0.0, b = 0.0, c = 0.0;
    for (long long i = start; i < n; i += stride) { a += (double)in[3 * i]; b += (double)in[3 * i + 1]; c += (double)in[3 * i + 2]; }
    __syncthreads();                    // a second call in the same block must not overwrite sm while the first result is still being read
    sm[0][threadIdx.x] = a; sm[1][threadIdx.x] = b; sm[2][threadIdx.x] = c;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) { sm[0][threadIdx.x] += sm[0][threadIdx.x + s]; sm[1][threadIdx.x] += sm[1][threadIdx.x + s]; sm[2][threadIdx.x] += sm[2][threadIdx.x + s]; }
        __syncthreads();
    }
    out[0] = sm[0][0]; out[1] = sm[1][0]; out[2] = sm[2][0];
}
// ---- what a slab knows about the other slabs of the decomposition (device-resident table, built at lbm_peer_attach*):
// every slab's adapter mailbox, IBM node mailbox and IBM stage counters, reached over NVLink peer mappings (own entries included).
constexpr int MAX_WORLD = 64;
struct AdpSlot { double s[3]; unsigned long long tag; };       // one slab's sums of (rho, rho|u|, |Pi|) for the step `tag`
struct SlabNet {
    int world, rank;
    AdpSlot* adp[MAX_WORLD];                    // [2][MAX_WORLD] slots per slab: parity of the step x source rank; nullptr = slab not mapped
    float* mail[MAX_WORLD];                     // IBM node mailboxes
    unsigned long long* ibm_stage[MAX_WORLD];   // [MAX_WORLD] IBM stage counters per slab, indexed by source rank
};

// The whole reduction in ONE launch: every block folds its share of the fp32 partials into an fp64 triple (stage[block]); the block
// that arrives last (ticket) adds the triples in block order — a fixed order, so the result does not depend on which block that is —
// and writes sums[3] and, on a single slab, avg[3] = sums / (NX*NY).  With `net` it also PUBLISHES this slab's sums for step `tag` into
// every slab's adapter mailbox (one 32-byte store per slab over NVLink, data first, tag after a system-wide fence): the send half of the
// device-side all-reduce of CM<2,OptimalAdapter>'s grid sums (macroscopics.cuh:161-177 is a single-GPU atomicAdd + symbol copy).
__global__ void __launch_bounds__(256) reduce_kernel(const float* partials, long long nparts, double* stage, unsigned* ticket, double* sums, float* avg,
                                                     double inv_n, int write_avg, const SlabNet* net, unsigned long long tag) {
    __shared__ bool last;
    double o[3];
    block_sum3(partials, nparts, (long long)blockIdx.x * 256 + threadIdx.x, (long long)gridDim.x * 256, o);
    if (threadIdx.x < 3) stage[3 * blockIdx.x + threadIdx.x] = o[threadIdx.x];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    block_sum3((const volatile double*)stage, (long long)gridDim.x, (long long)threadIdx.x, 256ll, o);
    if (threadIdx.x == 0) *ticket = 0u;
    if (threadIdx.x < 3) {
        sums[threadIdx.x] = o[threadIdx.x];
        if (write_avg) avg[threadIdx.x] = (float)(o[threadIdx.x] * inv_n);
    }
    if (net && (int)threadIdx.x < net->world) {
        AdpSlot* box = net->adp[threadIdx.x];
        if (box) {
            AdpSlot* dst = box + (tag & 1) * MAX_WORLD + net->rank;
            dst->s[0] = o[0]; dst->s[1] = o[1]; dst->s[2] = o[2];
            __threadfence_system();
            *(volatile unsigned long long*)&dst->tag = tag;
            __threadfence_system();
        }
    }
}
// the receive half: wait until every slab's sums for step `tag` have arrived in THIS slab's mailbox, add them in rank order
// (the same order on every slab, hence the same bits) and set the grid means the step kernels read
__global__ void __launch_bounds__(MAX_WORLD) adapter_collect_kernel(const SlabNet* net, unsigned long long tag, float* avg, double inv_n, int* timed_out) {
    const AdpSlot* box = net->adp[net->rank] + (tag & 1) * MAX_WORLD;
    const int r = threadIdx.x;
    if (r < net->world) {
        const volatile unsigned long long* tg = &box[r].tag;
        const long long t0 = clock64();
        while (*tg != tag) {
            if (clock64() - t0 > 20000000000ll) { *timed_out = 3; break; }      // ~10 s: a lost slab must not hang the GPU
            __nanosleep(100);
        }
        __threadfence_system();
    }
    __syncthreads();
    if (r < 3) {
        double a = 0.0;
        for (int k = 0; k < net->world; k++) a += *(const volatile double*)&box[k].s[r];
        avg[r] = (float)(a * inv_n);
    }
}
__global__ void sums_to_avg_kernel(const double* sums, float* avg, double inv_n) {
    if (threadIdx.x < 3) avg[threadIdx.x] = (float)(sums[threadIdx.x] * inv_n);
}

// post-stream populations of the interior neighbours that ZG_OUTFLOW / PRESSURE_OUTLET / corner nodes read
// (zeroGradientOutflow.cuh:44-57, pressureOutlet.cuh:12-26, regularizedBounceBack.cuh:136-147)
template <bool ODD>
__global__ void nbr_gather_kernel(const Params p, const long long* nbr_src, float* nbr_g, int count) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    long long node = nbr_src[k];
    int x = (int)(node % p.nx), yg = (int)(node / p.nx);
    float g[Q];
    pull<ODD>(p, x, yg - p.y0, g);
#pragma unroll
    for (int q = 0; q < Q; q++) nbr_g[(long long)k * Q + q] = g[q];
}

// ---- slab-to-slab step handshake for the peer-mapped mode (one thread each; flags live in the exporting slab's memory)
// flags[0] / flags[1]: last step completed by the lower / upper neighbour, written by that neighbour over NVLink.
__global__ void wait_neighbours_kernel(volatile unsigned long long* flags, int need_lo, int need_hi, unsigned long long t, int* timed_out) {
    const long long t0 = clock64();
    while ((need_lo && flags[0] < t) || (need_hi && flags[1] < t)) {
        if (clock64() - t0 > 20000000000ll) { *timed_out = 1; break; }       // ~10 s: a lost neighbour must not hang the GPU
        __nanosleep(200);
    }
}
__global__ void signal_neighbours_kernel(unsigned long long* lo_flag, unsigned long long* hi_flag, unsigned long long t) {
    __threadfence_system();
    if (lo_flag) *(volatile unsigned long long*)lo_flag = t;
    if (hi_flag) *(volatile unsigned long long*)hi_flag = t;
    __threadfence_system();
}

// one thread per segment: SEG_VEC / SEG_GENERAL / SEG_MIXED (allow_mixed = 0: a segment with any general cell is SEG_GENERAL as a
// whole — grids whose rows are not whole segments keep the shuffle-based odd kernel, which has no per-cell skipping)
__global__ void build_segmask_kernel(const Params p, uint8_t* mask, int allow_mixed) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)p.nsx * p.nyl) return;
    const int yl = (int)(i / p.nsx), sx = (int)(i - (long long)yl * p.nsx);
    if (p.force_plane) { mask[i] = SEG_GENERAL; return; }
    const int x1 = min(p.nx, (sx + 1) * SEG);
    int n = 0;
    for (int x = sx * SEG; x < x1; x++) n += cell_is_general(p, x, yl) ? 1 : 0;
    mask[i] = n == 0 ? SEG_VEC : ((n == x1 - sx * SEG || !allow_mixed) ? SEG_GENERAL : SEG_MIXED);
}
// the general cells of the SEG_MIXED segments, as local node ids (thrust::copy_if predicate)
struct mixed_general_cell {
    Params p; const uint8_t* mask;
    __device__ bool operator()(long long ln) const {
        const int yl = (int)(ln / p.nx), x = (int)(ln - (long long)yl * p.nx);
        return mask[(long long)yl * p.nsx + x / SEG] == SEG_MIXED && cell_is_general(p, x, yl);
    }
};

// ------------------------------------------------------------------ IBM
struct IbmData {
    int np, nnodes, ss;                 // markers, stencil nodes, stencil slots per marker (4 or 16)
    const long long* nodes;             // sorted unique global node ids
    const int* sten_idx;                // [np*ss] compact node index or -1
    const float* sten_w;                // [np*ss] delta4(dx)*delta4(dy)
    const int* row;                     // CSR node -> (marker, weight), markers ascending
    const int* csr_k; const float* csr_w;
    float* rho; float2* uprev; float2* lagF; float2* force;     // scratch + result
    const float2* utarget;              // [np] IBMBody::velocities, or nullptr = the reference's literal 0 (IBM_impl.cuh:15)
    // bodies across slab faces: the node states travel through a mailbox indexed by the GLOBAL stencil-node list
    // (IBM_MAIL floats per node), which every slab holds at the same offsets
    const int* mail_idx;                // [nnodes] mailbox slot of each node of this slab's (active) list
    float* mail;                        // this slab's mailbox
    const SlabNet* net;                 // peer-mapped slabs (nullptr = none): the owner of a node stores its state into THEIR mailboxes too
    unsigned long long need_mask;       // ranks that own a stencil node of the bodies this slab works on: their stage counters are awaited
    volatile unsigned long long* my_stage;   // this slab's IBM stage counters [MAX_WORLD], written by the slab of that rank
    int* timed_out;
};
constexpr int IBM_MAIL = 5;             // rho, u*_x, u*_y, F_x, F_y (d_force after reset_forces)

// state of one stencil node before the IBM iterations: interpolate_velocities_kernel's inputs (IBM_impl.cu:7-51) and the
// body force the result is accumulated onto (IBMManager.cuh:222-252)
template <bool ODD>
__device__ __forceinline__ void ibm_node_state(const Params& p, long long node, float& rho, float2& ustar, float2& F) {
    const int x = (int)(node % p.nx), yl = (int)(node / p.nx) - p.y0;
    float g[Q];
    pull<ODD>(p, x, yl, g);
    const long long ln = (long long)yl * p.nx + x;
    const int bc = p.flags ? (p.flags[ln] & FLAG_MASK) : 0;
    if (bc) apply_bc(p, bc, g, x, p.y0 + yl);
    const Moments m = moments(g);
    rho = m.rho;
    ustar = make_float2(m.ux, m.uy);                        // the uncorrected u* (IBMManager.cuh:227)
    F = p.force_plane ? p.force_plane[ln] : make_float2(p.fx, p.fy);     // d_force after reset_forces, before accumulation
}

// One direct-forcing iteration in two halves, shared by the one-block kernels and the many-block kernels below.
// Marker half: interpolate_velocities_kernel<2> (IBM_impl.cu:7-51) + compute_lagrangian_kernel (IBM_impl.cuh:9-26)
__device__ __forceinline__ void ibm_marker_pass(const IbmData& d, int k, bool clip) {
    float rho = 0.f, ux = 0.f, uy = 0.f;
    for (int s = 0; s < d.ss; s++) {
        int idx = d.sten_idx[k * d.ss + s];
        if (idx < 0) continue;
        float w = d.sten_w[k * d.ss + s];
        float2 u = d.uprev[idx];
        rho += w * d.rho[idx]; ux += w * u.x; uy += w * u.y;
    }
    const float2 ut = d.utarget ? d.utarget[k] : make_float2(0.0f, 0.0f);
    float Fx = 2.0f * rho * (ut.x - ux), Fy = 2.0f * rho * (ut.y - uy);
    if (clip) { Fx = Fx > 1e-8f ? Fx : 0.0f; Fy = Fy > 1e-8f ? Fy : 0.0f; }
    d.lagF[k] = make_float2(Fx, Fy);
}
// Node half: spread_forces_kernel<2> (IBM_impl.cu:122-154, as a gather), correct_velocities_kernel + accumulate_forces_kernel (IBM_impl.cuh:30-68)
__device__ __forceinline__ void ibm_node_pass(const IbmData& d, int i, bool clip) {
    float fx = 0.f, fy = 0.f;
    for (int e = d.row[i]; e < d.row[i + 1]; e++) { float2 F = d.lagF[d.csr_k[e]]; float w = d.csr_w[e]; fx += w * F.x; fy += w * F.y; }
    float r2 = 2.0f * d.rho[i];
    float2 u = d.uprev[i];
    float cux = u.x + fx / r2, cuy = u.y + fy / r2;
    if (clip) { cux = ((double)cux > 1e-8) ? cux : 0.0f; cuy = ((double)cuy > 1e-8) ? cuy : 0.0f; }
    d.uprev[i] = make_float2(cux, cuy);
    float2 F = d.force[i];
    d.force[i] = make_float2(F.x + fx, F.y + fy);
}

// 3 iterations of interpolate -> Lagrangian force -> spread -> correct -> accumulate over d.rho / d.uprev / d.force (one block)
__device__ __forceinline__ void ibm_iterations(const Params& p, const IbmData& d) {
    const bool clip = (p.quirks & QK_D7) != 0;
    for (int iter = 0; iter < 3; iter++) {                  // ITER_MAX (IBMManager.cuh:8)
        for (int k = threadIdx.x; k < d.np; k += blockDim.x) ibm_marker_pass(d, k, clip);
        __syncthreads();
        for (int i = threadIdx.x; i < d.nnodes; i += blockDim.x) ibm_node_pass(d, i, clip);
        __syncthreads();
    }
}

// The same as seven launches of many blocks, for bodies too large for one block to be quick (tens of thousands of markers):
// node states, then 3 x (marker half, node half).  Same per-marker / per-node arithmetic, hence the same bits.
template <bool ODD>
__global__ void __launch_bounds__(256) ibm_state_kernel(const Params p, const IbmData d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.nnodes) return;
    float rho; float2 us, F;
    ibm_node_state<ODD>(p, d.nodes[i], rho, us, F);
    d.rho[i] = rho; d.uprev[i] = us; d.force[i] = F;
}
__global__ void __launch_bounds__(256) ibm_markers_kernel(const Params p, const IbmData d) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < d.np) ibm_marker_pass(d, k, (p.quirks & QK_D7) != 0);
}
__global__ void __launch_bounds__(256) ibm_nodes_kernel(const Params p, const IbmData d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < d.nnodes) ibm_node_pass(d, i, (p.quirks & QK_D7) != 0);
}

// IBMManager<2>::multi_direct (src/IBM/IBMManager.cuh:222-252) as ONE launch working only on the nodes under
// marker stencils: interpolate_velocities_kernel<2> (IBM_impl.cu:7-51), compute_lagrangian_kernel
// (IBM_impl.cuh:9-26), spread_forces_kernel<2> (IBM_impl.cu:122-154; gather over a node<-marker CSR instead of
// atomicAdd, so the sum order is fixed), correct_velocities_kernel + accumulate_forces_kernel (IBM_impl.cuh:30-68).
template <bool ODD>
__global__ void __launch_bounds__(1024) ibm_kernel(const Params p, const IbmData d) {
    for (int i = threadIdx.x; i < d.nnodes; i += blockDim.x) {
        float rho; float2 us, F;
        ibm_node_state<ODD>(p, d.nodes[i], rho, us, F);
        d.rho[i] = rho; d.uprev[i] = us; d.force[i] = F;
    }
    __syncthreads();
    ibm_iterations(p, d);
}

// Bodies whose stencils cross a slab face.  Stage 1, on every slab: the states of the stencil nodes
// THIS slab owns go into mailbox `out` (its own, or a caller's buffer that is then all-reduced) and, peer-mapped, into the
// neighbours' mailboxes over NVLink, followed by the stage counter t.
template <bool ODD>
__global__ void __launch_bounds__(1024) ibm_gather_kernel(const Params p, const IbmData d, float* out, unsigned long long t) {
    for (int i = threadIdx.x; i < d.nnodes; i += blockDim.x) {
        const long long node = d.nodes[i];
        const int yg = (int)(node / p.nx);
        if (yg < p.y0 || yg >= p.y0 + p.nyl) continue;
        float rho; float2 us, F;
        ibm_node_state<ODD>(p, node, rho, us, F);
        const float v[IBM_MAIL] = {rho, us.x, us.y, F.x, F.y};
        const long long o = (long long)d.mail_idx[i] * IBM_MAIL;
#pragma unroll
        for (int c = 0; c < IBM_MAIL; c++) out[o + c] = v[c];
        if (d.net)
            for (int r = 0; r < d.net->world; r++) {
                float* pm = d.net->mail[r];
                if (!pm || r == d.net->rank) continue;
#pragma unroll
                for (int c = 0; c < IBM_MAIL; c++) pm[o + c] = v[c];
            }
    }
    if (d.net) {
        __threadfence_system();
        __syncthreads();
        const int r = threadIdx.x;
        if (r < d.net->world && r != d.net->rank && d.net->ibm_stage[r]) {
            *(volatile unsigned long long*)(d.net->ibm_stage[r] + d.net->rank) = t;
            __threadfence_system();
        }
    }
}
// Stage 2, on every slab that owns part of a body: wait for the neighbours' stage counters (peer-mapped only), read the
// complete node states from the mailbox and run the iterations redundantly — every slab of a body computes the same bits.
__global__ void __launch_bounds__(1024) ibm_solve_kernel(const Params p, const IbmData d, unsigned long long t) {
    if (d.need_mask && threadIdx.x < MAX_WORLD && ((d.need_mask >> threadIdx.x) & 1ull)) {
        const long long t0 = clock64();
        while (d.my_stage[threadIdx.x] < t) {
            if (clock64() - t0 > 20000000000ll) { *d.timed_out = 2; break; }
            __nanosleep(200);
        }
        __threadfence_system();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < d.nnodes; i += blockDim.x) {
        const float* m = d.mail + (long long)d.mail_idx[i] * IBM_MAIL;
        d.rho[i] = __ldcv(m);
        d.uprev[i] = make_float2(__ldcv(m + 1), __ldcv(m + 2));
        d.force[i] = make_float2(__ldcv(m + 3), __ldcv(m + 4));
    }
    __syncthreads();
    ibm_iterations(p, d);
}

// delta4 / kernel2D — reference src/IBM/IBMUtils.cuh:23-43
__device__ __forceinline__ float delta4(float r) {
    float rabs = fabsf(r);
    if (rabs < 1.0f) return 0.125f * (3.0f - 2.0f * rabs + sqrtf(1.0f + 4.0f * rabs - 4.0f * r * r));
    else if (rabs < 2.0f) return 0.125f * (5.0f - 2.0f * rabs - sqrtf(-7.0f + 12.0f * rabs - 4.0f * r * r));
    return 0.0f;
}

// per marker: the stencil node ids (or -1 outside the domain) and weights, slot order i (x) outer, j (y) inner
__global__ void ibm_stencil_kernel(const float* pts, int np, int nx, int ny, int lo, int w, long long* sten_node, float* sten_w) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= np) return;
    float px = pts[2 * k], py = pts[2 * k + 1];
    float gx = floorf(px), gy = floorf(py);
    for (int i = 0; i < w; i++)
        for (int j = 0; j < w; j++) {
            int nxx = (int)(gx + (i + lo)), nyy = (int)(gy + (j + lo));
            int s = k * w * w + i * w + j;
            if (nxx >= nx || nxx < 0 || nyy >= ny || nyy < 0) { sten_node[s] = -1; sten_w[s] = 0.f; continue; }
            float dx = px - nxx, dy = py - nyy;
            // contraction off: the weights are compared bit-for-bit with the CPU restatement
            sten_w[s] = __fmul_rn(delta4(dx), delta4(dy));
            sten_node[s] = (long long)nyy * nx + nxx;
        }
}

// ------------------------------------------------------------------ init / readback
// init_kernel + init_node (src/core/init/init.cuh:10-43) with equilibrium_node's fp32/fp64 mix (equilibrium.cu:5-39).
// Writes the layout expected by an odd first step: A[opp q][x] = f_q(x); both rest planes; both ring generations.
__device__ __forceinline__ void store_initial(const Params& p, int x, int yl, const float f[Q]) {
    const long long row = rowoff(p, yl);
    p.A0[0][row + x] = f[0]; p.A0[1][row + x] = f[0];
#pragma unroll
    for (int q = 1; q < Q; q++) p.A[opp(q)][row + x] = f[q];
    int e = edge_index(p, x, p.y0 + yl);
    if (e >= 0)
        for (int gen = 0; gen < 2; gen++)
#pragma unroll
            for (int q = 0; q < Q; q++) p.ring[((long long)gen * p.perim + e) * Q + q] = f[q];
}

__device__ __forceinline__ void feq_reference(float rho, float ux, float uy, float f[Q]) {
    float u_dot_u = ux * ux + uy * uy;
    float cs = 1.0f / sqrtf(3.0f);
    float cs2 = cs * cs, cs4 = cs2 * cs2;
#pragma unroll
    for (int q = 0; q < Q; q++) {
        float cu = __fadd_rn(__fmul_rn((float)cx(q), ux), __fmul_rn((float)cy(q), uy));
        double cud = (double)cu;
        double br = 1 + 0.5 * (cud * cud) / cs4 - 0.5 * u_dot_u / cs2 + 1.0 * cu / cs2;
        f[q] = (float)((double)(wq(q) * rho) * br);
    }
}

__global__ void __launch_bounds__(BX) init_fields_kernel(const Params p, const float* rho, const float2* u) {
    const int x = blockIdx.x * BX + threadIdx.x, yl = blockIdx.y;
    if (x >= p.nx) return;
    const long long ln = (long long)yl * p.nx + x;
    float f[Q];
    float2 uu = u[ln];
    feq_reference(rho[ln], uu.x, uu.y, f);
    store_initial(p, x, yl, f);
    if (p.rho_out) { p.rho_out[ln] = rho[ln]; p.u_out[ln] = uu; }
}

// TaylorGreenInit::operator() — reference src/scenarios/taylorGreen/taylorGreenFunctors.cuh:25-47
__global__ void __launch_bounds__(BX) init_taylor_green_kernel(const Params p, float nu, float u0) {
    const int xi = blockIdx.x * BX + threadIdx.x, yl = blockIdx.y;
    if (xi >= p.nx) return;
    const float x = xi + 0.5f, y = (p.y0 + yl) + 0.5f;
    const float kx = (float)(2.0 * 3.14159265358979323846 / p.nx), ky = (float)(2.0 * 3.14159265358979323846 / p.ny);
    float ux = -u0 * sqrtf(ky / kx) * cosf(kx * x) * sinf(ky * y);
    float uy = u0 * sqrtf(kx / ky) * sinf(kx * x) * cosf(ky * y);
    float P = -0.25f * u0 * u0 * ((ky / kx) * cosf(2 * kx * x) + (kx / ky) * cosf(2 * ky * y));
    float rho = 1.0f + 3.0f * P;
    float f[Q];
    feq_reference(rho, ux, uy, f);
    store_initial(p, xi, yl, f);
    if (p.rho_out) { const long long ln = (long long)yl * p.nx + xi; p.rho_out[ln] = rho; p.u_out[ln] = make_float2(ux, uy); }
}

// from AoS post-collision populations f[node*9+q] (what the reference's d_f holds after collide()); timestep even
__global__ void __launch_bounds__(BX) set_populations_kernel(const Params p, const float* f_aos, const float* fb_aos) {
    const int x = blockIdx.x * BX + threadIdx.x, yl = blockIdx.y;
    if (x >= p.nx) return;
    const long long ln = (long long)yl * p.nx + x, row = rowoff(p, yl);
    const float* f = f_aos + ln * Q; const float* fb = fb_aos + ln * Q;
    // the next step is t+1: it reads the rest plane / ring generation (t+1)&1, which in the reference is f_back
    const int g_next = (p.t + 1) & 1;
    p.A0[g_next][row + x] = fb[0];
    p.A0[g_next ^ 1][row + x] = f[0];
    if (p.A0[0] == p.A0[1]) p.A0[0][row + x] = f[0];
    for (int q = 1; q < Q; q++) p.A[opp(q)][row + x] = f[q];
    int e = edge_index(p, x, p.y0 + yl);
    if (e >= 0)
        for (int q = 0; q < Q; q++) {
            p.ring[((long long)g_next * p.perim + e) * Q + q] = fb[q];
            p.ring[((long long)(g_next ^ 1) * p.perim + e) * Q + q] = f[q];
        }
}

// post-collision population f*_q(x) of the last step, wherever the AA phase left it
template <bool LAST_ODD>
__device__ __forceinline__ void read_post_collision(const Params& p, int x, int yl, float f[Q]) {
    const int gen = p.t & 1;                  // p.t = last completed step
    const long long row = rowoff(p, yl);
    f[0] = p.A0[gen][row + x];
    if (!LAST_ODD) {
#pragma unroll
        for (int q = 1; q < Q; q++) f[q] = p.A[opp(q)][row + x];
    } else {
        int e = edge_index(p, x, p.y0 + yl);
#pragma unroll
        for (int q = 1; q < Q; q++) {
            int xd = x, yd = yl;
            bool ok = hop(p, xd, yd, cx(q), cy(q));
            f[q] = ok ? row_ptr(p, q, yd)[xd] : p.ring[((long long)gen * p.perim + e) * Q + q];
        }
    }
}

template <bool LAST_ODD>
__global__ void __launch_bounds__(BX) get_populations_kernel(const Params p, float* f_aos) {
    const int x = blockIdx.x * BX + threadIdx.x, yl = blockIdx.y;
    if (x >= p.nx) return;
    float f[Q];
    read_post_collision<LAST_ODD>(p, x, yl, f);
    const long long ln = (long long)yl * p.nx + x;
#pragma unroll
    for (int q = 0; q < Q; q++) f_aos[ln * Q + q] = f[q];
}

template <bool LAST_ODD>
__global__ void __launch_bounds__(BX) mass_kernel(const Params p, double* out) {
    const int x = blockIdx.x * BX + threadIdx.x, yl = blockIdx.y;
    double v = 0.0;
    if (x < p.nx) {
        float f[Q];
        read_post_collision<LAST_ODD>(p, x, yl, f);
#pragma unroll
        for (int q = 0; q < Q; q++) v += (double)f[q];
    }
    __shared__ double sm[BX];
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int s = BX / 2; s > 0; s >>= 1) { if ((int)threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s]; __syncthreads(); }
    if (threadIdx.x == 0) atomicAdd(out, sm[0]);
}

// rho and u of the LAST step rebuilt from its post-collision populations, for callers that ask for the macroscopic fields of a
// step that was enqueued without them.  Collision conserves mass and changes momentum by a known amount dJ (Guo forcing:
// dJ = F for BGK, CM and MRT; with the reference's swapped MRT force rows, A-D2, dJ_y = S5 F_y / 2 - F_x (1 - S5 / 2)), so
// rho = sum f*, u* = (sum f* c - dJ) / rho, u = u* + F / (2 rho): the reference's d_rho / d_u up to fp32 round-off.
template <bool LAST_ODD>
__global__ void __launch_bounds__(BX) recover_macros_kernel(const Params p) {
    const int x = blockIdx.x * BX + threadIdx.x, yl = blockIdx.y;
    if (x >= p.nx) return;
    float f[Q];
    read_post_collision<LAST_ODD>(p, x, yl, f);
    const Moments m = moments(f);
    const long long ln = (long long)yl * p.nx + x;
    float Fx = p.fx, Fy = p.fy;
    if (p.force_plane) { const float2 F = p.force_plane[ln]; Fx = F.x; Fy = F.y; }
    if (p.flags && (p.flags[ln] & FLAG_IBM)) {
        const int k = find_sorted(p.ibm_nodes, p.ibm_count, (long long)(p.y0 + yl) * p.nx + x);
        if (k >= 0) { const float2 F = p.ibm_force[k]; Fx = F.x; Fy = F.y; }
    }
    if (p.t == 0) Fx = Fy = 0.0f;            // the initial state: no collision has added momentum yet, d_u = the Init functor's u
    float dJx = Fx, dJy = Fy;
    if (p.coll == C_MRT && (p.quirks & QK_D2)) dJy = 0.5f * p.S[5] * Fy - Fx * (1.0f - 0.5f * p.S[5]);
    const float jx = m.ux * m.rho - dJx, jy = m.uy * m.rho - dJy;
    const float h = 0.5f * m.inv_rho;
    p.rho_out[ln] = m.rho;
    p.u_out[ln] = make_float2(fmaf(Fx, h, jx * m.inv_rho), fmaf(Fy, h, jy * m.inv_rho));
}

// ------------------------------------------------------------------ validation reductions (deterministic, fp64)
// TaylorGreenValidation::operator() — reference src/scenarios/taylorGreen/taylorGreenFunctors.cuh:66-81 (u0 = its u_max/SCALE)
__device__ __forceinline__ float2 taylor_green_analytic(int xi, int yi, int nx, int ny, float nu, float u0, float t) {
    const float x = xi + 0.5f, y = yi + 0.5f;
    const float kx = 2.0f * (float)3.14159265358979323846 / nx, ky = 2.0f * (float)3.14159265358979323846 / ny;
    const float td = 1.0f / (nu * (kx * kx + ky * ky));
    const float decay = expf(-t / td);
    return make_float2(-u0 * sqrtf(ky / kx) * cosf(kx * x) * sinf(ky * y) * decay, u0 * sqrtf(kx / ky) * sinf(kx * x) * cosf(ky * y) * decay);
}

__device__ __forceinline__ void block_sum2(double a, double b, double* out) {
    __shared__ double sm[2][256];
    sm[0][threadIdx.x] = a; sm[1][threadIdx.x] = b;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) { sm[0][threadIdx.x] += sm[0][threadIdx.x + s]; sm[1][threadIdx.x] += sm[1][threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = sm[0][0]; out[1] = sm[1][0]; }
}

// the two sums of Scenario::compute_error (taylorGreenScenario.cuh:66-87): sum |u - u_ref|^2 and sum |u_ref|^2 over this slab;
// u_ref from a device field (AoS, slab rows) or, TG = true, the analytic Taylor-Green field evaluated in place
template <bool TG>
__global__ void __launch_bounds__(256) error_sums_kernel(const float2* u, const float2* ref, int nx, int ny, int y0, long long n, float nu, float u0, float t, double* stage) {
    double e = 0.0, r = 0.0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        float2 a;
        if (TG) { const int yl = (int)(i / nx); a = taylor_green_analytic((int)(i - (long long)yl * nx), y0 + yl, nx, ny, nu, u0, t); }
        else a = ref[i];
        const float2 s = u[i];
        const float dx = s.x - a.x, dy = s.y - a.y;
        e += (double)(dx * dx + dy * dy);
        r += (double)(a.x * a.x + a.y * a.y);
    }
    block_sum2(e, r, stage + 2 * blockIdx.x);
}
__global__ void __launch_bounds__(256) error_sums_final_kernel(const double* stage, int n, double* out) {
    double e = 0.0, r = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) { e += stage[2 * i]; r += stage[2 * i + 1]; }
    block_sum2(e, r, out);
}
// u at a list of global nodes (the samples of a centre-line validation: LidDrivenScenario::compute_error reads 2 x 17 of them from
// h_u, lidDrivenCavityScenario.cuh:103-135); nodes other slabs own are left untouched
__global__ void sample_velocity_kernel(const float2* u, const long long* nodes, int n, long long first, long long nloc, float2* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long ln = nodes[i] - first;
    if (ln >= 0 && ln < nloc) out[i] = u[ln];
}
// mean of u over x for every row of the slab (the inner loop of PoiseuilleScenario::compute_error, poiseuilleScenario.cuh:63-70)
__global__ void __launch_bounds__(256) row_mean_kernel(const float2* u, int nx, double* mean_ux, double* mean_uy) {
    const float2* row = u + (long long)blockIdx.x * nx;
    double a = 0.0, b = 0.0;
    for (int x = threadIdx.x; x < nx; x += 256) { const float2 v = row[x]; a += (double)v.x; b += (double)v.y; }
    double o[2];
    block_sum2(a, b, o);
    if (threadIdx.x == 0) { mean_ux[blockIdx.x] = o[0] / nx; mean_uy[blockIdx.x] = o[1] / nx; }
}

}  // namespace lbm
