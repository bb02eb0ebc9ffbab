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
    const float h = 0.5f / s.m.rho;
    s.ux = s.m.ux + s.Fx * h; s.uy = s.m.uy + s.Fy * h;
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
        for (int i = 0; i < BX / 32; i++) { s0 += sm[0][i]; s1 += sm[1][i]; s2 += sm[2][i]; }
        out[0] = s0; out[1] = s1; out[2] = s2;
    }
}

// The fused step: one launch = one reference time step (src/main.cu:96-114) for every node of the slab.
template <int COLL, bool ODD, bool GENERAL>
__global__ void __launch_bounds__(BX) step_kernel(const Params p) {
    const int x = blockIdx.x * BX + threadIdx.x;
    const int yl = blockIdx.y;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (x < p.nx) {
        NodeState s;
        node_pre_collision<ODD, GENERAL>(p, x, yl, s);
        if (p.rho_out) {
            const long long ln = (long long)yl * p.nx + x;
            p.rho_out[ln] = s.m.rho;
            p.u_out[ln] = make_float2(s.ux, s.uy);
        }
        if (COLL == C_CMOPT) { s0 = s.m.rho; s1 = s.m.rho * sqrtf(s.ux * s.ux + s.uy * s.uy); s2 = pi_norm(s.m); }
        if (COLL == C_BGK) collide_bgk(p, s.g, s.m.rho, s.ux, s.uy, s.Fx, s.Fy);
        else if (COLL == C_MRT) collide_mrt(p, s.g, s.m.rho, s.ux, s.uy, s.Fx, s.Fy);
        else if (COLL == C_CM) collide_cm<false>(p, s.g, s.ux, s.uy, s.Fx, s.Fy);
        else collide_cm<true>(p, s.g, s.ux, s.uy, s.Fx, s.Fy);
        push<ODD>(p, x, yl, s.e, s.g);
    }
    if (COLL == C_CMOPT && p.partials) block_partials(s0, s1, s2, p.partials + 3 * ((long long)blockIdx.y * gridDim.x + blockIdx.x));
}

// moments pre-pass for LBM_ADAPTER_EXACT: the grid sums of the CURRENT post-stream state, before any cell collides
template <bool ODD>
__global__ void __launch_bounds__(BX) moments_kernel(const Params p) {
    const int x = blockIdx.x * BX + threadIdx.x;
    const int yl = blockIdx.y;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (x < p.nx) {
        NodeState s;
        node_pre_collision<ODD, true>(p, x, yl, s);
        s0 = s.m.rho; s1 = s.m.rho * sqrtf(s.ux * s.ux + s.uy * s.uy); s2 = pi_norm(s.m);
    }
    block_partials(s0, s1, s2, p.partials + 3 * ((long long)blockIdx.y * gridDim.x + blockIdx.x));
}

// deterministic final reduction of the block partials: sums[3] (fp64) and avg[3] = sums / (NX*NY)
__global__ void reduce_partials_kernel(const float* partials, long long nblocks, double* sums, float* avg, double inv_n, int write_avg) {
    __shared__ double sm[3][256];
    double a = 0.0, b = 0.0, c = 0.0;
    for (long long i = threadIdx.x; i < nblocks; i += 256) { a += partials[3 * i]; b += partials[3 * i + 1]; c += partials[3 * i + 2]; }
    sm[0][threadIdx.x] = a; sm[1][threadIdx.x] = b; sm[2][threadIdx.x] = c;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) { sm[0][threadIdx.x] += sm[0][threadIdx.x + s]; sm[1][threadIdx.x] += sm[1][threadIdx.x + s]; sm[2][threadIdx.x] += sm[2][threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x < 3) {
        sums[threadIdx.x] = sm[threadIdx.x][0];
        if (write_avg) avg[threadIdx.x] = (float)(sm[threadIdx.x][0] * inv_n);
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

// ------------------------------------------------------------------ IBM
struct IbmData {
    int np, nnodes, ss;                 // markers, stencil nodes, stencil slots per marker (4 or 16)
    const long long* nodes;             // sorted unique global node ids
    const int* sten_idx;                // [np*ss] compact node index or -1
    const float* sten_w;                // [np*ss] delta4(dx)*delta4(dy)
    const int* row;                     // CSR node -> (marker, weight), markers ascending
    const int* csr_k; const float* csr_w;
    float* rho; float2* uprev; float2* lagF; float2* force;     // scratch + result
};

// IBMManager<2>::multi_direct (src/IBM/IBMManager.cuh:222-252) as ONE launch working only on the nodes under
// marker stencils: interpolate_velocities_kernel<2> (IBM_impl.cu:7-51), compute_lagrangian_kernel
// (IBM_impl.cuh:9-26), spread_forces_kernel<2> (IBM_impl.cu:122-154; gather over a node<-marker CSR instead of
// atomicAdd, so the sum order is fixed), correct_velocities_kernel + accumulate_forces_kernel (IBM_impl.cuh:30-68).
template <bool ODD>
__global__ void __launch_bounds__(1024) ibm_kernel(const Params p, const IbmData d) {
    const bool clip = (p.quirks & QK_D7) != 0;
    for (int i = threadIdx.x; i < d.nnodes; i += blockDim.x) {
        long long node = d.nodes[i];
        int x = (int)(node % p.nx), yl = (int)(node / p.nx) - p.y0;
        float g[Q];
        pull<ODD>(p, x, yl, g);
        const long long ln = (long long)yl * p.nx + x;
        int bc = p.flags ? (p.flags[ln] & FLAG_MASK) : 0;
        if (bc) apply_bc(p, bc, g, x, p.y0 + yl);
        Moments m = moments(g);
        d.rho[i] = m.rho;
        d.uprev[i] = make_float2(m.ux, m.uy);               // the uncorrected u* (IBMManager.cuh:227)
        float2 F = p.force_plane ? p.force_plane[ln] : make_float2(p.fx, p.fy);
        d.force[i] = F;                                     // d_force after reset_forces, before accumulation
    }
    __syncthreads();
    for (int iter = 0; iter < 3; iter++) {                  // ITER_MAX (IBMManager.cuh:8)
        for (int k = threadIdx.x; k < d.np; k += blockDim.x) {
            float rho = 0.f, ux = 0.f, uy = 0.f;
            for (int s = 0; s < d.ss; s++) {
                int idx = d.sten_idx[k * d.ss + s];
                if (idx < 0) continue;
                float w = d.sten_w[k * d.ss + s];
                float2 u = d.uprev[idx];
                rho += w * d.rho[idx]; ux += w * u.x; uy += w * u.y;
            }
            float Fx = 2.0f * rho * (0.0f - ux), Fy = 2.0f * rho * (0.0f - uy);
            if (clip) { Fx = Fx > 1e-8f ? Fx : 0.0f; Fy = Fy > 1e-8f ? Fy : 0.0f; }
            d.lagF[k] = make_float2(Fx, Fy);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < d.nnodes; i += blockDim.x) {
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
        __syncthreads();
    }
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
            f[q] = ok ? p.A[q][rowoff(p, yd) + xd] : p.ring[((long long)gen * p.perim + e) * Q + q];
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

}  // namespace lbm
